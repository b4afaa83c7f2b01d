"""ctypes binding of libjaicov_b200.so (the C ABI declared in include/jaicov_b200.h).

There is no fallback: if the shared library is missing the import raises, and if no sm_100 device is usable every
computing call returns JAICOV_NOT_INITIALISED, which ``check`` turns into an exception.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libjaicov_b200.so')

OK, ERROR_FREE_ESTIMATION, INTERRUPT, SINGULAR_MATRIX, NO_CONVERGENCE = 0, 1, -1, -2, -4
NOT_INITIALISED, OUT_OF_MEMORY, ILLEGAL_ARGUMENT = -5, -7, -100
INVERT_NONE, INVERT_FULL, INVERT_PRE_ELIMINATION, INVERT_REDUCED = 0, 1, 2, 3
L2NORM, SIMULATION = 0, 1
SOLVER_AUTO, SOLVER_DENSE, SOLVER_STRUCTURED = 0, 1, 2
COL_UNSET, COL_FIXED = -1, 2147483647

EXPORTS = [
    'jaicov_default_options', 'jaicov_create', 'jaicov_destroy', 'jaicov_last_error', 'jaicov_device_count', 'jaicov_launch_count', 'jaicov_release_cached_memory', 'jaicov_nccl_unique_id', 'jaicov_dist_init',
    'jaicov_get_qxx_local', 'jaicov_shard_images',
    'jaicov_set_cameras', 'jaicov_set_images', 'jaicov_set_image_points', 'jaicov_set_object_points',
    'jaicov_set_scale_bars', 'jaicov_add_observed_group', 'jaicov_set_datum', 'jaicov_set_reduced_rows', 'jaicov_estimate', 'jaicov_iterate',
    'jaicov_get_stats', 'jaicov_get_values', 'jaicov_get_dx', 'jaicov_get_qxx_packed', 'jaicov_get_qxx_block',
    'jaicov_get_qxx_diag', 'jaicov_get_qxx_submatrix', 'jaicov_eval_residual_jacobian', 'jaicov_get_normal_equations', 'jaicov_omega',
    'jaicov_spd_solve_invert', 'jaicov_propagate_eo_transform', 'jaicov_dlt_batch', 'jaicov_gemm_tiles',
    'jaicov_normal_product', 'jaicov_get_preconditioner', 'jaicov_get_sweep_times', 'jaicov_set_image_dispersion', 'jaicov_set_gemm_digits', 'jaicov_get_device_bytes',
]


class Options(ctypes.Structure):
    _fields_ = [('invert_mode', ctypes.c_int32), ('estimation_type', ctypes.c_int32), ('max_iterations', ctypes.c_int32),
                ('use_centroid', ctypes.c_int32), ('apply_aposteriori', ctypes.c_int32), ('device', ctypes.c_int32),
                ('solver', ctypes.c_int32), ('n_devices', ctypes.c_int32),
                ('sigma2apriori', ctypes.c_double), ('damping_value', ctypes.c_double)]


class Stats(ctypes.Structure):
    _fields_ = [('status', ctypes.c_int32), ('iterations', ctypes.c_int32), ('iteration_step', ctypes.c_int32),
                ('n_unknowns', ctypes.c_int32), ('n_datum', ctypes.c_int32), ('n_observations', ctypes.c_int32),
                ('dof', ctypes.c_int32), ('solver_used', ctypes.c_int32),
                ('omega', ctypes.c_double), ('max_abs_dx', ctypes.c_double), ('sigma2apriori', ctypes.c_double),
                ('sigma2aposteriori', ctypes.c_double),
                ('ms_assembly', ctypes.c_double), ('ms_factor', ctypes.c_double), ('ms_solve', ctypes.c_double),
                ('ms_inverse', ctypes.c_double), ('ms_omega', ctypes.c_double), ('ms_total', ctypes.c_double)]


PROGRESS_CB = ctypes.CFUNCTYPE(None, ctypes.c_void_p, ctypes.c_int32, ctypes.c_double, ctypes.c_double)

_lib = None


def load():
    """Loads the shared library (raises OSError when it has not been built: run __graft_entry__.build())."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise OSError('%s is missing: build it with `python -c "import __graft_entry__ as g; g.build()"` '
                      '(jaicov_b200 has no CPU fallback)' % LIB_PATH)
    L = ctypes.CDLL(LIB_PATH)
    vp, i32, i64, dbl = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_double
    L.jaicov_default_options.argtypes = [ctypes.POINTER(Options)]
    L.jaicov_create.argtypes = [ctypes.POINTER(Options), ctypes.POINTER(vp)]
    L.jaicov_destroy.argtypes = [vp]
    L.jaicov_destroy.restype = None
    L.jaicov_last_error.argtypes = [vp]
    L.jaicov_last_error.restype = ctypes.c_char_p
    L.jaicov_device_count.argtypes = []
    L.jaicov_launch_count.argtypes = []
    L.jaicov_nccl_unique_id.argtypes = [vp]
    L.jaicov_shard_images.argtypes = [i32, vp, i32, i32, ctypes.POINTER(i32), ctypes.POINTER(i32)]
    L.jaicov_dist_init.argtypes = [vp, i32, i32, vp]
    L.jaicov_get_qxx_local.argtypes = [vp, vp, vp, i32, vp]
    L.jaicov_set_cameras.argtypes = [vp, i32, vp, vp, vp, vp, vp, vp, vp, vp]
    L.jaicov_set_images.argtypes = [vp, i32, vp, vp, vp, vp]
    L.jaicov_set_image_points.argtypes = [vp, i64, vp, vp, vp, vp]
    L.jaicov_set_object_points.argtypes = [vp, i32, vp, vp, vp]
    L.jaicov_set_scale_bars.argtypes = [vp, i32, vp, vp, vp, vp]
    L.jaicov_add_observed_group.argtypes = [vp, i32, vp, vp, vp, vp, vp, vp]
    L.jaicov_set_datum.argtypes = [vp, vp, i32, i32]
    L.jaicov_set_reduced_rows.argtypes = [vp, i32]
    L.jaicov_estimate.argtypes = [vp, vp, vp, vp]
    L.jaicov_iterate.argtypes = [vp, i32, i32]
    L.jaicov_get_stats.argtypes = [vp, ctypes.POINTER(Stats)]
    L.jaicov_get_values.argtypes = [vp, vp, vp, vp, vp]
    L.jaicov_get_dx.argtypes = [vp, vp]
    L.jaicov_get_qxx_packed.argtypes = [vp, vp]
    L.jaicov_get_qxx_block.argtypes = [vp, i32, i32, i32, i32, vp, i64]
    L.jaicov_get_qxx_diag.argtypes = [vp, vp]
    L.jaicov_get_qxx_submatrix.argtypes = [vp, i32, vp, dbl, vp]
    L.jaicov_eval_residual_jacobian.argtypes = [vp, i32, vp, vp, vp]
    L.jaicov_release_cached_memory.restype = ctypes.c_int64
    L.jaicov_release_cached_memory.argtypes = []
    L.jaicov_propagate_eo_transform.argtypes = [vp, i32, vp, vp, vp, dbl, vp, vp]
    L.jaicov_dlt_batch.argtypes = [i32, i32, vp, vp, vp, vp, i32, vp, i32, vp, vp, vp]
    L.jaicov_get_normal_equations.argtypes = [vp, vp, vp]
    L.jaicov_omega.argtypes = [vp, vp, ctypes.POINTER(dbl)]
    L.jaicov_normal_product.argtypes = [vp, i32, vp, vp, vp, ctypes.POINTER(dbl)]
    L.jaicov_get_preconditioner.argtypes = [vp, vp]
    L.jaicov_set_gemm_digits.argtypes = [i32]
    L.jaicov_get_device_bytes.argtypes = [vp, vp]
    L.jaicov_set_image_dispersion.argtypes = [vp, i32, i64, vp]
    L.jaicov_get_sweep_times.argtypes = [vp, ctypes.POINTER(dbl), ctypes.POINTER(dbl), ctypes.POINTER(dbl)]
    L.jaicov_spd_solve_invert.argtypes = [i32, i64, vp, i32, vp, i32, ctypes.POINTER(dbl), ctypes.POINTER(dbl)]
    L.jaicov_gemm_tiles.argtypes = [i32] * 5 + [i64, dbl, dbl, vp, i64, vp, i64, vp, i64, i32, i32, i32, vp]
    for name in EXPORTS:
        if name not in ('jaicov_destroy', 'jaicov_last_error', 'jaicov_launch_count', 'jaicov_release_cached_memory'):
            getattr(L, name).restype = ctypes.c_int32
    L.jaicov_launch_count.restype = ctypes.c_int64
    L.jaicov_release_cached_memory.restype = ctypes.c_int64       # bytes: tens of GB
    _lib = L
    return L


class JaicovError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__('jaicov_b200 error %d: %s' % (code, msg))
        self.code = code


def _p(a):
    return None if a is None or a.size == 0 else a.ctypes.data


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def dlt_batch(pt_ptr, xy, xyz, io, restrictions=(), max_iterations=5000, device=0):
    """Batched direct linear transformation (jaicov_dlt_batch).  Returns (out (n_img, 20), status (n_img), passes (n_img))."""
    L = load()
    pt_ptr = np.ascontiguousarray(pt_ptr, dtype=np.int64)
    n_img = pt_ptr.size - 1
    xy, xyz, io = _f64(xy).reshape(-1), _f64(xyz).reshape(-1), _f64(io).reshape(-1)
    restr = _i32(list(restrictions))
    out = np.zeros((max(n_img, 0), 20))
    status = np.zeros(max(n_img, 0), np.int32)
    passes = np.zeros(max(n_img, 0), np.int32)
    rc = L.jaicov_dlt_batch(device, n_img, _p(pt_ptr), _p(xy), _p(xyz), _p(io), restr.size, _p(restr), int(max_iterations),
                            _p(out), _p(status), _p(passes))
    if rc != OK:
        raise JaicovError(rc, 'jaicov_dlt_batch failed (no sm_100 device, or illegal argument)')
    return out, status, passes


class Session:
    """One adjustment on one device: thin object wrapper around a ``jaicov_handle``.

    ``flat`` is a dict of flat arrays in the layout of include/jaicov_b200.h (see ``set_problem``)."""

    def __init__(self, invert_mode=INVERT_FULL, estimation_type=L2NORM, max_iterations=5000, use_centroid=True,
                 apply_aposteriori=True, device=0, sigma2apriori=1.0, damping_value=0.0, solver=SOLVER_AUTO, n_devices=1):
        self.L = load()
        self.opt = Options()
        self.L.jaicov_default_options(ctypes.byref(self.opt))
        self.opt.invert_mode = invert_mode
        self.opt.estimation_type = estimation_type
        self.opt.max_iterations = max_iterations
        self.opt.use_centroid = int(use_centroid)
        self.opt.apply_aposteriori = int(apply_aposteriori)
        self.opt.device = device
        self.opt.solver = solver
        self.opt.n_devices = n_devices
        self.opt.sigma2apriori = sigma2apriori
        self.opt.damping_value = damping_value
        self.h = ctypes.c_void_p()
        rc = self.L.jaicov_create(ctypes.byref(self.opt), ctypes.byref(self.h))
        if rc != OK:
            raise JaicovError(rc, 'jaicov_create failed (unsupported option?)')
        self._keep = []
        self._interrupt = ctypes.c_int32(0)
        self.n = 0
        self.flat = None

    def close(self):
        if self.h:
            self.L.jaicov_destroy(self.h)
            self.h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, rc, allowed=(OK,)):
        if rc not in allowed:
            raise JaicovError(rc, (self.L.jaicov_last_error(self.h) or b'').decode())
        return rc

    def set_problem(self, f):
        """f: dict with io_val, io_col, r0, coef_ptr, coef_type, coef_order, coef_val, coef_col, cam_of_img, eo_val,
        eo_col, pt_ptr, obj_idx, xy, var, rho, xyz, pt_col, is_datum, [bar_a, bar_b, bar_len, bar_var], [groups],
        free_flags, n_unknowns, n_observations."""
        L, h = self.L, self.h
        self.flat = f
        io_val, io_col, r0 = _f64(f['io_val']), _i32(f['io_col']), _f64(f['r0'])
        cp, ct, co = _i32(f['coef_ptr']), _i32(f['coef_type']), _i32(f['coef_order'])
        cv, cc = _f64(f['coef_val']), _i32(f['coef_col'])
        self.check(L.jaicov_set_cameras(h, r0.size, _p(io_val), _p(io_col), _p(r0), _p(cp), _p(ct), _p(co), _p(cv), _p(cc)))
        coi, ev, ec = _i32(f['cam_of_img']), _f64(f['eo_val']), _i32(f['eo_col'])
        pp = np.ascontiguousarray(f['pt_ptr'], dtype=np.int64)
        self.check(L.jaicov_set_images(h, coi.size, _p(coi), _p(ev), _p(ec), _p(pp)))
        oi, xy, var, rho = _i32(f['obj_idx']), _f64(f['xy']), _f64(f['var']), _f64(f['rho'])
        self.check(L.jaicov_set_image_points(h, oi.size, _p(oi), _p(xy), _p(var), _p(rho)))
        xyz, pc = _f64(f['xyz']), _i32(f['pt_col'])
        dat = np.ascontiguousarray(f['is_datum'], dtype=np.uint8)
        self.check(L.jaicov_set_object_points(h, xyz.size // 3, _p(xyz), _p(pc), _p(dat)))
        if 'bar_a' in f and len(f['bar_a']):
            ba, bb, bl, bv = _i32(f['bar_a']), _i32(f['bar_b']), _f64(f['bar_len']), _f64(f['bar_var'])
            self.check(L.jaicov_set_scale_bars(h, ba.size, _p(ba), _p(bb), _p(bl), _p(bv)))
        for g in f.get('groups', []):
            k, ix, cm, ob = _i32(g['kind']), _i32(g['index']), _i32(g['comp']), _f64(g['obs'])
            var_g = None if g.get('var') is None else _f64(g['var'])
            sig = None if g.get('sigma') is None else _f64(g['sigma'])
            self.check(L.jaicov_add_observed_group(h, k.size, _p(k), _p(ix), _p(cm), _p(ob), _p(var_g), _p(sig)))
        for (img_i, sig) in f.get('img_sigma', []):
            sig = _f64(sig)
            nr = int(round((-1 + (1 + 8 * sig.size) ** 0.5) / 2))
            self.check(L.jaicov_set_image_dispersion(h, int(img_i), nr, _p(sig)))
        ff = _i32(f['free_flags'])
        self.check(L.jaicov_set_datum(h, _p(ff), int(f['n_unknowns']), int(f['n_observations'])))
        self.n = int(f['n_unknowns']) + int(ff.sum())
        self.n_qxx = self.n
        if f.get('reduced_rows') is not None:
            self.check(L.jaicov_set_reduced_rows(h, int(f['reduced_rows'])))
            if self.opt.invert_mode in (INVERT_REDUCED, INVERT_PRE_ELIMINATION):
                self.n_qxx = min(self.n, int(f['reduced_rows']))
        self.shapes = dict(xyz=xyz.size, io=io_val.size, coef=cv.size, eo=ev.size, m=oi.size,
                           ncoef_max=int(np.max(np.diff(cp))) if cp.size > 1 else 0)

    def dist_init(self, rank, world, nccl_id: bytes):
        """Joins the NCCL communicator of a multi-GPU adjustment (call before set_problem)."""
        buf = ctypes.create_string_buffer(bytes(nccl_id), 128)
        self.check(self.L.jaicov_dist_init(self.h, rank, world, buf))
        self.rank, self.world = rank, world

    def qxx_local(self, out=None):
        """(first reference column of every owned 128-wide tile, list of (np - e_i) x 128 blocks) of a distributed
        handle; ``out``: optional (pinned) float64 buffer of at least ``qxx_local_size()`` elements."""
        nt = ctypes.c_int32(0)
        self.check(self.L.jaicov_get_qxx_local(self.h, ctypes.byref(nt), None, 0, None))
        cols = np.zeros(max(nt.value, 1), np.int32)
        self.check(self.L.jaicov_get_qxx_local(self.h, ctypes.byref(nt), cols.ctypes.data, nt.value, None))
        cols = cols[:nt.value]
        u, d = int(self.flat['n_unknowns']), self.n - int(self.flat['n_unknowns'])
        npad = (max(u, 1) + 127) // 128 * 128
        sizes = [(npad - (int(c) - d)) * 128 for c in cols]
        total = int(sum(sizes))
        if out is None:
            out = np.empty(max(total, 1))
        self.check(self.L.jaicov_get_qxx_local(self.h, ctypes.byref(nt), None, 0, out.ctypes.data if total else None))
        blocks, off = [], 0
        for sz in sizes:
            blocks.append(out[off:off + sz].reshape(-1, 128))
            off += sz
        return cols, blocks

    def interrupt(self):
        """Ask a running estimate() to stop at its next check (BundleAdjustment.interrupt, BA:1455-1457); callable from the
        progress callback or from another thread."""
        self._interrupt.value = 1

    def estimate(self, progress=None):
        cb = PROGRESS_CB(lambda user, st, a, b: progress(st, a, b)) if progress else None
        rc = self.L.jaicov_estimate(self.h, ctypes.cast(cb, ctypes.c_void_p) if cb else None, None, ctypes.byref(self._interrupt))
        if rc in (NOT_INITIALISED, OUT_OF_MEMORY, ILLEGAL_ARGUMENT):
            self.check(rc)
        return rc

    def iterate(self, final_pass=False, apply_update=True):
        rc = self.L.jaicov_iterate(self.h, int(final_pass), int(apply_update))
        if rc in (NOT_INITIALISED, OUT_OF_MEMORY, ILLEGAL_ARGUMENT):
            self.check(rc)
        return rc

    def stats(self):
        s = Stats()
        self.check(self.L.jaicov_get_stats(self.h, ctypes.byref(s)))
        return s

    def values(self):
        sh = self.shapes
        xyz, io, coef, eo = np.zeros(sh['xyz']), np.zeros(sh['io']), np.zeros(sh['coef']), np.zeros(sh['eo'])
        self.check(self.L.jaicov_get_values(self.h, _p(xyz), _p(io), _p(coef), _p(eo)))
        return xyz, io, coef, eo

    def dx(self):
        out = np.zeros(self.n)
        self.check(self.L.jaicov_get_dx(self.h, _p(out)))
        return out

    def qxx_packed(self, out=None):
        if out is None:
            out = np.empty(self.n_qxx * (self.n_qxx + 1) // 2)
        self.check(self.L.jaicov_get_qxx_packed(self.h, out.ctypes.data))
        return out

    def qxx_block(self, r0, r1, c0, c1):
        out = np.empty((r1 - r0, c1 - c0))
        self.check(self.L.jaicov_get_qxx_block(self.h, r0, r1, c0, c1, out.ctypes.data, c1 - c0))
        return out

    def qxx_submatrix(self, idx, scale=1.0):
        idx = _i32(idx)
        out = np.empty((idx.size, idx.size))
        self.check(self.L.jaicov_get_qxx_submatrix(self.h, idx.size, _p(idx), float(scale), _p(out)))
        return out

    def propagate_eo_transform(self, point, src_image, trg_image, sigma2, covariance=True):
        """Transformed coordinates (R, 3) and the packed upper covariance sigma2 * J Qxx J' of the triples
        (point, source image, target image); see jaicov_propagate_eo_transform."""
        point, src_image, trg_image = _i32(point), _i32(src_image), _i32(trg_image)
        n = point.size
        xyz = np.empty((n, 3))
        cov = np.empty(3 * n * (3 * n + 1) // 2) if covariance else None
        self.check(self.L.jaicov_propagate_eo_transform(self.h, n, _p(point), _p(src_image), _p(trg_image), float(sigma2),
                                                        _p(xyz), _p(cov) if covariance else None))
        return xyz, cov

    def qxx_diag(self):
        out = np.empty(self.n_qxx)
        self.check(self.L.jaicov_get_qxx_diag(self.h, out.ctypes.data))
        return out

    def eval_residual_jacobian(self):
        sh = self.shapes
        ns = 12 + sh['ncoef_max']
        a, w, p = np.zeros((sh['m'], 2, ns)), np.zeros((sh['m'], 2)), np.zeros((sh['m'], 3))
        self.check(self.L.jaicov_eval_residual_jacobian(self.h, ns, _p(a), _p(w), _p(p)))
        return a, w, p

    def normal_equations(self):
        N, rhs = np.zeros(self.n * (self.n + 1) // 2), np.zeros(self.n)
        self.check(self.L.jaicov_get_normal_equations(self.h, _p(N), _p(rhs)))
        return N, rhs

    def omega(self, dx):
        dx = _f64(dx)
        out = ctypes.c_double(0.0)
        self.check(self.L.jaicov_omega(self.h, _p(dx), ctypes.byref(out)))
        return out.value

    def normal_product(self, X=None, want_rhs=True):
        """(Y, rhs, w'Pw): Y = K X' for the rows of X ([nvec][u+d]) with the bordered normal matrix, matrix-free from the
        observations (jaicov_normal_product); rhs = [0; A'Pw]."""
        X = np.zeros((0, self.n)) if X is None else np.ascontiguousarray(np.atleast_2d(np.asarray(X, np.float64)))
        Y = np.zeros_like(X)
        rhs = np.zeros(self.n) if want_rhs else None
        wpw = ctypes.c_double(0.0)
        self.check(self.L.jaicov_normal_product(self.h, X.shape[0], _p(X) if X.size else None, _p(Y) if Y.size else None,
                                                _p(rhs) if want_rhs else None, ctypes.byref(wpw) if want_rhs else None))
        return Y, rhs, wpw.value

    def sweep_times(self):
        """Device ms of the by-image, by-point and Omega sweeps of the last pass."""
        a, b, c = ctypes.c_double(0), ctypes.c_double(0), ctypes.c_double(0)
        self.check(self.L.jaicov_get_sweep_times(self.h, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c)))
        return a.value, b.value, c.value

    def device_bytes(self):
        """Bytes of (system matrix, second square, inverse column tiles, panel staging) on the device (jaicov_get_device_bytes)."""
        out = np.zeros(4, np.int64)
        self.check(self.L.jaicov_get_device_bytes(self.h, _p(out)))
        return [int(v) for v in out]

    def preconditioner(self):
        out = np.empty(self.n)
        self.check(self.L.jaicov_get_preconditioner(self.h, _p(out)))
        return out


def set_gemm_digits(digits):
    """Process-wide arithmetic of the big tile products (jaicov_set_gemm_digits): 0 = FP64 DMMA tiles, 4..8 = int8 digit products."""
    return load().jaicov_set_gemm_digits(int(digits))


def shard_images(pt_ptr, world, rank):
    """Image range [begin, end) of `rank` (host-only; same rule the library applies internally)."""
    L = load()
    pp = np.ascontiguousarray(pt_ptr, dtype=np.int64)
    b, e = ctypes.c_int32(0), ctypes.c_int32(0)
    rc = L.jaicov_shard_images(pp.size - 1, pp.ctypes.data, world, rank, ctypes.byref(b), ctypes.byref(e))
    if rc != OK:
        raise JaicovError(rc, 'jaicov_shard_images')
    return b.value, e.value


def nccl_unique_id() -> bytes:
    L = load()
    buf = ctypes.create_string_buffer(128)
    rc = L.jaicov_nccl_unique_id(buf)
    if rc != OK:
        raise JaicovError(rc, 'jaicov_nccl_unique_id')
    return buf.raw


def spd_solve_invert(a, b=None, invert=True, device=0):
    """Level-1 seam: SPD solve / inverse on the device (see jaicov_spd_solve_invert).  Returns (inverse|None, x|None, ms)."""
    L = load()
    a = np.array(a, dtype=np.float64, order='C')
    n = a.shape[0]
    nrhs = 0
    if b is not None:
        b = np.array(np.atleast_2d(b), dtype=np.float64, order='C')
        nrhs = b.shape[0]
    f, i = ctypes.c_double(0), ctypes.c_double(0)
    rc = L.jaicov_spd_solve_invert(device, n, a.ctypes.data, nrhs, _p(b), int(invert), ctypes.byref(f), ctypes.byref(i))
    if rc not in (OK,):
        raise JaicovError(rc, 'jaicov_spd_solve_invert')
    return (a if invert else None), b, (f.value, i.value)


def gemm_tiles(A, B, C, a_layout=0, b_layout=0, alpha=1.0, beta=0.0, tri_out=False, kmode=0, reps=1, device=0):
    """Stage access to the tensor-core tile product (jaicov_gemm_tiles): C <- alpha op(A) op(B)' + beta C on 128 x 128 tiles.
    A: (128 mt, K) if a_layout == 0 else (K, 128 mt); B likewise with 128 nt; C: (128 mt, 128 nt).  Returns (C, ms)."""
    L = load()
    A = np.ascontiguousarray(A, dtype=np.float64)
    B = np.ascontiguousarray(B, dtype=np.float64)
    C = np.array(C, dtype=np.float64, order='C')
    Mr, K = (A.shape if a_layout == 0 else A.shape[::-1])
    Nr = B.shape[0] if b_layout == 0 else B.shape[1]
    ms = ctypes.c_double(0)
    rc = L.jaicov_gemm_tiles(device, a_layout, b_layout, Mr // 128, Nr // 128, K, alpha, beta, A.ctypes.data, A.shape[1], B.ctypes.data,
                             B.shape[1], C.ctypes.data, C.shape[1], int(bool(tri_out)), kmode, reps, ctypes.byref(ms))
    if rc != OK:
        raise JaicovError(rc, 'jaicov_gemm_tiles')
    return C, ms.value
