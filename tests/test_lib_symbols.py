"""The C-ABI library loads and exports every symbol include/jaicov_b200.h declares; without a GPU it fails loudly."""
import ctypes
import os
import re

import numpy as np
import pytest

import bundle_adjustment_b200 as ba

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, 'include', 'jaicov_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(jaicov_[a-z0-9_]+)\s*\(', src)) - {'jaicov_progress_cb'})


def test_exports_match_header(built):
    L = ba._lib.load()
    names = declared_symbols()
    assert len(names) >= 31
    for n in names:
        assert hasattr(L, n), n
    assert sorted(ba._lib.EXPORTS) == names


def test_every_entry_point_is_documented_for_the_integrator():
    """INTEGRATION.md section 2 maps every C entry point to the reference interface it replaces (or says it has none)."""
    doc = open(os.path.join(ROOT, 'INTEGRATION.md')).read()
    missing = [n for n in declared_symbols() if n not in doc]
    assert not missing, missing


def test_struct_sizes(built):
    assert ctypes.sizeof(ba._lib.Options) == 8 * 4 + 2 * 8
    assert ctypes.sizeof(ba._lib.Stats) == 8 * 4 + 10 * 8


def test_option_validation(built):
    L = ba._lib.load()
    opt = ba._lib.Options()
    assert L.jaicov_default_options(ctypes.byref(opt)) == 0
    assert (opt.invert_mode, opt.max_iterations, opt.use_centroid, opt.apply_aposteriori) == (1, 5000, 1, 1)
    h = ctypes.c_void_p()
    opt.invert_mode = 7                                 # unknown mode -> refused
    assert L.jaicov_create(ctypes.byref(opt), ctypes.byref(h)) == ba._lib.ILLEGAL_ARGUMENT
    opt.invert_mode = ba._lib.INVERT_FULL
    opt.damping_value = -0.1                            # negative damping -> refused
    assert L.jaicov_create(ctypes.byref(opt), ctypes.byref(h)) == ba._lib.ILLEGAL_ARGUMENT
    opt.damping_value = 0.1
    assert L.jaicov_create(ctypes.byref(opt), ctypes.byref(h)) == 0
    L.jaicov_destroy(h)


def _has_device():
    try:
        return ba._lib.load().jaicov_device_count() > 0
    except Exception:
        return False


@pytest.mark.skipif(_has_device(), reason='a B200 is visible: the calls below would compute')
def test_no_cpu_fallback(built):
    """No B200 -> computing entry points fail with NOT_INITIALISED instead of falling back to the CPU."""
    a = np.eye(4)
    with pytest.raises(ba.JaicovError) as e:
        ba.spd_solve_invert(a)
    assert e.value.code == ba._lib.NOT_INITIALISED
    s = ba.Session()
    with pytest.raises(ba.JaicovError):
        s.set_problem(dict(io_val=np.zeros(3), io_col=np.array([7, 8, 9]), r0=np.ones(1), coef_ptr=[0, 0], coef_type=[],
                           coef_order=[], coef_val=[], coef_col=[], cam_of_img=[0], eo_val=np.zeros(6),
                           eo_col=np.arange(10, 16), pt_ptr=[0, 1], obj_idx=[0], xy=np.zeros(2), var=np.ones(2),
                           rho=np.zeros(1), xyz=np.zeros(3), pt_col=[0, 1, 2], is_datum=[1], free_flags=[0] * 7,
                           n_unknowns=16, n_observations=2))
        s.iterate(final_pass=True)


@pytest.mark.skipif(_has_device(), reason='a B200 is visible: the calls below would compute')
def test_no_cpu_fallback_for_the_round2_entry_points(built):
    """The single-process multi-GPU handle, the matrix-free product, the dense image dispersion and the device queries fail loudly
    (NOT_INITIALISED) without devices -- and the handle still closes cleanly."""
    from bundle_adjustment_b200.workloads import flat_problem, synthetic_scene
    sc = synthetic_scene(2, images=4, targets=20)[0]
    m2 = 2 * len(sc['cameras'][0]['images'][1]['obj'])
    sc['cameras'][0]['images'][1]['dispersion'] = np.full(m2 * (m2 + 1) // 2, 0.0)      # content irrelevant here
    adj, flat = flat_problem(sc)
    assert [i for i, _s in flat['img_sigma']] == [1]
    for nd in (1, 2):
        s = ba.Session(sigma2apriori=adj.getVarianceFactorApriori(), n_devices=nd)
        s.set_problem(flat)
        with pytest.raises(ba.JaicovError) as e:
            s.estimate()
        assert e.value.code == ba._lib.NOT_INITIALISED and 'no CPU path' in str(e.value)
        with pytest.raises(ba.JaicovError) as e:
            s.normal_product(np.zeros((1, s.n)))
        assert e.value.code == ba._lib.NOT_INITIALISED
        with pytest.raises(ba.JaicovError):
            s.qxx_packed()
        assert s.device_bytes() == [0, 0, 0, 0]
        s.close()
    assert ba._lib.set_gemm_digits(-1) in (-1, 0, 8)             # query only: nothing decided without a device
    with pytest.raises(ba.JaicovError):
        ba.Session(n_devices=-1)


@pytest.mark.skipif(_has_device(), reason='a B200 is visible: the calls below would compute')
def test_no_cpu_fallback_for_the_widened_entry_points(built):
    """jaicov_dlt_batch and the covariance propagation have no CPU route either."""
    pt_ptr = np.array([0, 6])
    xy = np.zeros((6, 2))
    xyz = np.arange(18.0).reshape(6, 3)
    with pytest.raises(ba.JaicovError) as e:
        ba._lib.dlt_batch(pt_ptr, xy, xyz, np.array([[30.0, 0.0, 0.0]]))
    assert e.value.code == ba._lib.NOT_INITIALISED
    s = ba.Session()
    with pytest.raises(ba.JaicovError) as e:
        s.propagate_eo_transform([0], [0], [0], 1.0)          # no problem set, no cofactor matrix
    assert e.value.code in (ba._lib.NOT_INITIALISED, ba._lib.ILLEGAL_ARGUMENT)
    # illegal arguments are refused before any device work
    L = ba._lib.load()
    assert L.jaicov_dlt_batch(0, 1, None, None, None, None, 0, None, 10, None, None, None) == ba._lib.ILLEGAL_ARGUMENT
    assert L.jaicov_release_cached_memory() == 0


def test_host_mirror_of_the_widened_classes():
    """Pure host behaviour of the mirrors: Image.get (camera/Image.java:77-79), DLTCoefficients order
    (dlt/DLTCoefficients.java:38-64), transform() without a device-resident Qxx fails loudly."""
    pts = ba.ObjectCoordinateArray(['a', 'b', 'c'], np.arange(9.0).reshape(3, 3))
    cam = ba.Camera(1, 10.0)
    img = cam.add(7)
    img.addAll(pts, [0, 2], [[1.0, 2.0], [3.0, 4.0]], [[0.01, 0.01]] * 2)
    img.add(pts[1], 5.0, 6.0, 0.01, 0.01)
    assert img.get(pts[0]) == (1.0, 2.0) and img.get(pts[2]) == (3.0, 4.0) and img.get(pts[1]) == (5.0, 6.0)
    other = ba.ObjectCoordinate('z', 0, 0, 0)
    assert img.get(other) is None
    coef = ba.DLTCoefficients(img)
    ids = [int(p.getParameterType()) for p in coef]
    assert ids == [611, 612, 613, 614, 621, 622, 623, 624, 631, 632, 633, 111, 112, 113, 251, 252, 253, 261, 262, 263]
    assert coef.getReference() is img
    t = ba.CoordinateTransformationExteriorOrientation.getInstance()
    assert t is ba.CoordinateTransformationExteriorOrientation.getInstance()
    with pytest.raises(ba.JaicovError):
        t.transform([pts[0]], {img: [img]}, 1.0, ba.UpperSymmPackMatrix(3, np.zeros(6)))
    RT = ba.DirectLinearTransformation.RestrictionType
    assert [r.value for r in RT] == [0, 1, 2, 3, 4, 5] and RT.FIXED_PRINCIPAL_POINT_Y.name == 'FIXED_PRINCIPAL_POINT_Y'


def test_null_arrays_are_refused_not_dereferenced(built):
    """The ABI never aborts: a missing array with a positive count is JAICOV_ILLEGAL_ARGUMENT with a message."""
    L = ba._lib.load()
    opt = ba._lib.Options()
    L.jaicov_default_options(ctypes.byref(opt))
    h = ctypes.c_void_p()
    assert L.jaicov_create(ctypes.byref(opt), ctypes.byref(h)) == 0
    I = ba._lib.ILLEGAL_ARGUMENT
    assert L.jaicov_set_cameras(h, 1, None, None, None, None, None, None, None, None) == I
    assert L.jaicov_set_cameras(h, 0, None, None, None, None, None, None, None, None) == 0
    assert L.jaicov_set_images(h, 1, None, None, None, None) == I
    assert L.jaicov_set_image_points(h, ctypes.c_int64(3), None, None, None, None) == I
    assert L.jaicov_set_object_points(h, 2, None, None, None) == I
    assert L.jaicov_set_scale_bars(h, 1, None, None, None, None) == I
    one = np.ones(1)
    assert L.jaicov_add_observed_group(h, 1, None, None, None, None, one.ctypes.data, None) == I
    assert L.jaicov_set_datum(h, None, 3, 3) == I
    L.jaicov_last_error.restype = ctypes.c_char_p
    assert b'null array' in L.jaicov_last_error(h)
    L.jaicov_destroy(h)


def test_struct_offsets_used_by_the_java_binding(built):
    """INTEGRATION.md section 3 addresses jaicov_options / jaicov_stats by byte offset from Java (MemorySegment.set/get):
    those offsets are part of the ABI."""
    O, S = ba._lib.Options, ba._lib.Stats
    assert [getattr(O, f).offset for f in ('invert_mode', 'estimation_type', 'max_iterations', 'use_centroid', 'apply_aposteriori',
                                            'device', 'solver', 'sigma2apriori', 'damping_value')] == [0, 4, 8, 12, 16, 20, 24, 32, 40]
    assert [getattr(S, f).offset for f in ('status', 'iterations', 'iteration_step', 'n_unknowns', 'n_datum', 'n_observations', 'dof',
                                            'solver_used', 'omega', 'max_abs_dx', 'sigma2apriori', 'sigma2aposteriori')] == \
        [0, 4, 8, 12, 16, 20, 24, 28, 32, 40, 48, 56]
    import re
    doc = open(os.path.join(ROOT, 'INTEGRATION.md')).read()
    assert 'opt.set(JAVA_DOUBLE, 32, sigma2apriori)' in doc and 'opt.set(JAVA_DOUBLE, 40, dampingValue)' in doc
    assert re.search(r'st\.get\(JAVA_DOUBLE, 32\).*st\.get\(JAVA_DOUBLE, 40\).*st\.get\(JAVA_INT, 8\)', doc)
