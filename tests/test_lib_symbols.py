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


@pytest.mark.skipif(ba._lib.load().jaicov_device_count() > 0, reason='only meaningful without a GPU')
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
