"""Test/bench helper: raw scenes (the plain-data description of an adjustment before bookkeeping).

``example_scene()`` loads tests/golden/example_scene.npz (BASELINE.json configs[0], generated from the
reference's bundled AICON report by tests/golden/make_example_fixture.py).
"""
import os

import numpy as np

_GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def example_scene():
    z = np.load(os.path.join(_GOLDEN, 'example_scene.npz'))
    images = []
    ptr = z['pt_ptr']
    for i in range(z['eo_val'].shape[0]):
        a, b = int(ptr[i]), int(ptr[i + 1])
        images.append({'eo_val': z['eo_val'][i].copy(), 'eo_fixed': np.zeros(6, bool), 'obj': z['obj'][a:b].copy(),
                       'xy': z['xy'][a:b].copy(), 'sigma': z['sigma'][a:b].copy(), 'rho': np.zeros(b - a)})
    cam = {'r0': float(z['r0']), 'io_val': z['io_val'].copy(), 'io_fixed': z['io_fixed'].copy(),
           'coefs': [(int(t), int(o), float(v), bool(f)) for t, o, v, f in
                     zip(z['coef_type'], z['coef_order'], z['coef_val'], z['coef_fixed'])],
           'images': images}
    npt = z['point_xyz'].shape[0]
    return {
        'points': {'xyz': z['point_xyz'].copy(), 'fixed': np.zeros((npt, 3), bool), 'datum': z['point_datum'].copy(),
                   'names': [str(s) for s in z['point_name']]},
        'cameras': [cam],
        'scale_bars': [(int(a), int(b), float(l), float(s)) for a, b, l, s in
                       zip(z['bar_a'], z['bar_b'], z['bar_len'], z['bar_sigma'])],
        'observed_groups': [],
    }


# the synthetic networks live in the package (bench.py uses them too)
from bundle_adjustment_b200.workloads import (CONFIGS, IO_TRUTH, _angles_from_rotation, _rotation, project,  # noqa: E402,F401
                                              random_scene, synthetic_scene)
