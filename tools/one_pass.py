#!/usr/bin/env python
"""One final pass of one configuration on one GPU (for ncu: no warm-up, no repetition).  python tools/one_pass.py <config> <dense|structured> [passes]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bundle_adjustment_b200 as ba  # noqa: E402
from bundle_adjustment_b200.workloads import flat_problem, synthetic_scene  # noqa: E402

config = int(sys.argv[1]) if len(sys.argv) > 1 else 5
solver = sys.argv[2] if len(sys.argv) > 2 else 'dense'
passes = int(sys.argv[3]) if len(sys.argv) > 3 else 1
adj, flat = flat_problem(synthetic_scene(config)[0])
s = ba.Session(sigma2apriori=adj.getVarianceFactorApriori(), solver={'dense': ba._lib.SOLVER_DENSE, 'structured': ba._lib.SOLVER_STRUCTURED}[solver])
s.set_problem(flat)
for _ in range(passes):
    assert s.iterate(final_pass=True, apply_update=False) == 0
st = s.stats()
print('config %d %s: %.2f ms (assembly %.2f, factor %.2f, solve %.2f, inverse %.2f, omega %.2f), sweeps by-image/by-point/omega %s ms'
      % (config, solver, st.ms_total, st.ms_assembly, st.ms_factor, st.ms_solve, st.ms_inverse, st.ms_omega, ['%.3f' % t for t in s.sweep_times()]))
