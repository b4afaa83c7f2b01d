"""Size-independent verification of one final pass, on the device, without any CPU reference.

The CPU oracle (test infrastructure) stops at n ~ 2e4; the benchmarked configuration has n = 63 020 and, on several GPUs, a
factorisation spread over many panels.  What can be checked at ANY size are the defining identities of the quantities the
reference returns (BundleAdjustment.java:228-355 with MathExtension.solve, MathExtension.java:338-366):

  * solution      K [lambda; dx] = [0; n]            (the bordered system dspsv solves; rows 0..d-1 are  B dx = 0)
  * cofactors     K Qxx e_c = e_c                    (what dsptri returns), on sampled columns c
  * Omega         v'Pv = w'Pw - 2 n'dx + dx'N dx     (getOmega, :472-491: v = w - A dx; exact for ANY dx, so it tests the Omega sweep
                                                      itself -- the shorter w'Pw - n'dx holds only for an exact solution and is
                                                      first-order sensitive to the rounding of dx far from convergence)

K x is evaluated by ``jaicov_normal_product``: matrix-free, straight from the observations with the model code of the
assembly -- it never reads the assembled matrix, its factor or Qxx, so the check is independent of the solver route (dense,
structured) and of the storage layout (one GPU, block-cyclic panels).  Residuals are measured on the Jacobi-preconditioned
system the library factors (V K V, unit diagonal; NormalEquationSystem.java:82-91), where they are comparable across
parameters of very different units.  Used by bench.py after the timed region at every N and by tests/ at full config sizes.
"""
import numpy as np


def sample_columns(n, d, world=1, panel=1024, extra=4, seed=0):
    """First and last column, the border, both sides of a few panel boundaries of the distributed factorisation (128-tile and
    `panel`-column boundaries in internal numbering = reference column - d), one column inside every rank's first tiles, and a
    few random ones."""
    cols = {0, d, n - 1, max(d, n - 129)}
    u = n - d
    for b in (128, panel, 2 * panel, (u // 2) // panel * panel, (u - 1) // panel * panel):
        for off in (-1, 0):
            if 0 <= b + off < u:
                cols.add(d + b + off)
    for r in range(world):                      # inverse tiles are dealt out in snake order: tile r belongs to rank r
        if 128 * r + 7 < u:
            cols.add(d + 128 * r + 7)
    rng = np.random.default_rng(seed)
    cols.update(int(c) for c in rng.integers(0, n, size=extra))
    return np.array(sorted(c for c in cols if 0 <= c < n), dtype=np.int64)


def check_pass(sess, columns=None, reduce_sum=None, omega=None, with_cofactors=True, values_updated=False):
    """Residuals of the last final pass of ``sess`` (a ``Session``).  ``reduce_sum(array) -> array`` sums a host array over
    the ranks of a distributed handle (in place or not); every rank must call this function with the same arguments.
    ``values_updated``: the pass applied its dx to the parameters (``estimate()``, ``iterate(apply_update=True)``) -- the right-hand
    side at the new values is no longer the one dx was solved for, so the solution residual is not defined and is skipped (the
    cofactor columns and, at convergence, Omega still are).
    Returns a dict of scaled residuals (all should be ~ eps * cond of the preconditioned system)."""
    n = sess.n
    u = int(sess.flat['n_unknowns'])
    d = n - u
    sol = sess.dx()                             # [lambda; dx]
    V = sess.preconditioner()
    cols = np.zeros(0, np.int64)
    Q = np.zeros((0, n))
    if with_cofactors:
        cols = sample_columns(n, d) if columns is None else np.asarray(columns, np.int64)
        Q = np.empty((cols.size, n))
        for i, c in enumerate(cols):
            Q[i] = sess.qxx_block(0, n, int(c), int(c) + 1)[:, 0]
        if reduce_sum is not None:
            Q = reduce_sum(Q)
    X = np.vstack([sol[None, :], Q])
    Y, rhs, wpw = sess.normal_product(X)
    out = {'n': int(n), 'd': int(d), 'columns': [int(c) for c in cols]}
    # solution
    r = V * (Y[0] - rhs)
    ytil = sol / V
    out['solve_residual'] = (None if values_updated else
                             float(np.max(np.abs(r)) / (np.max(np.abs(V * rhs)) + np.max(np.abs(ytil[d:])) + 1e-300)))
    dxn = float(np.max(np.abs(sol[d:]))) if u else 0.0
    out['datum_residual'] = float(np.max(np.abs(Y[0][:d])) / (dxn + 1e-300)) if d else 0.0
    # cofactor columns
    worst = 0.0
    per = []
    for i, c in enumerate(cols):
        res = V * Y[1 + i] / V[c]
        res[c] -= 1.0
        e = float(np.max(np.abs(res)))
        per.append(e)
        worst = max(worst, e)
    out['cofactor_residual'] = worst
    out['cofactor_residual_per_column'] = per
    # Omega
    if omega is not None:
        ident = wpw - 2.0 * float(rhs @ sol) + float(sol @ Y[0]) - 2.0 * float(sol[:d] @ Y[0][:d])   # dx'N dx = s'Ks - 2 lambda'(B dx)
        out['omega'] = float(omega)
        out['omega_identity'] = float(ident)
        out['omega_rel_diff'] = float(abs(omega - ident) / max(abs(omega), 1e-300))
    return out


def assert_ok(chk, tol_solve=1e-8, tol_cofactor=1e-8, tol_omega=1e-8):
    bad = []
    if chk['solve_residual'] is not None and not chk['solve_residual'] <= tol_solve:
        bad.append('solve_residual %.3g' % chk['solve_residual'])
    if not chk['datum_residual'] <= tol_solve:
        bad.append('datum_residual %.3g' % chk['datum_residual'])
    if chk['columns'] and not chk['cofactor_residual'] <= tol_cofactor:
        bad.append('cofactor_residual %.3g' % chk['cofactor_residual'])
    if 'omega_rel_diff' in chk and not chk['omega_rel_diff'] <= tol_omega:
        bad.append('omega_rel_diff %.3g' % chk['omega_rel_diff'])
    if bad:
        raise AssertionError('verification of the pass failed: ' + ', '.join(bad) + ' ' + repr({k: v for k, v in chk.items() if k != 'cofactor_residual_per_column'}))
    return chk
