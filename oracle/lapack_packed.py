"""ORACLE (test infrastructure, not the product): the LAPACK routines the reference calls.

The reference's factor/solve/invert is third-party binary code:
``MathExtension.solve(UpperSymmPackMatrix, ...)`` -> ``dspsv`` + ``dsptri`` (MathExtension.java:338-366),
``MathExtension.inv(UpperSPDPackMatrix)`` -> ``dpptrf`` + ``dpptri`` (MathExtension.java:304-324), executed by
``net.sourceforge.f2j:arpack_combined_all:0.1`` (F2J translation of reference LAPACK) behind
``com.github.fommil.netlib:core:1.1.2`` and ``mtj:1.0.4`` (all binary jars under JAICOV/lib, no source).
The same reference-LAPACK algorithms are exported by scipy's bundled OpenBLAS; they are reached here
through the function pointers in ``scipy.linalg.cython_lapack.__pyx_capi__``.
"""
from __future__ import annotations

import ctypes

import numpy as np
from scipy.linalg import cython_lapack

_c_int_p = ctypes.POINTER(ctypes.c_int)
_c_dbl_p = ctypes.POINTER(ctypes.c_double)

ctypes.pythonapi.PyCapsule_GetName.restype = ctypes.c_char_p
ctypes.pythonapi.PyCapsule_GetName.argtypes = [ctypes.py_object]
ctypes.pythonapi.PyCapsule_GetPointer.restype = ctypes.c_void_p
ctypes.pythonapi.PyCapsule_GetPointer.argtypes = [ctypes.py_object, ctypes.c_char_p]


def _fn(name, argtypes):
    cap = cython_lapack.__pyx_capi__[name]
    ptr = ctypes.pythonapi.PyCapsule_GetPointer(cap, ctypes.pythonapi.PyCapsule_GetName(cap))
    return ctypes.CFUNCTYPE(None, *argtypes)(ptr)


_dspsv = _fn('dspsv', [ctypes.c_char_p, _c_int_p, _c_int_p, _c_dbl_p, _c_int_p, _c_dbl_p, _c_int_p, _c_int_p])
_dsptrf = _fn('dsptrf', [ctypes.c_char_p, _c_int_p, _c_dbl_p, _c_int_p, _c_int_p])
_dsptri = _fn('dsptri', [ctypes.c_char_p, _c_int_p, _c_dbl_p, _c_int_p, _c_dbl_p, _c_int_p])
_dpptrf = _fn('dpptrf', [ctypes.c_char_p, _c_int_p, _c_dbl_p, _c_int_p])
_dpptri = _fn('dpptri', [ctypes.c_char_p, _c_int_p, _c_dbl_p, _c_int_p])


def _p(a):
    return a.ctypes.data_as(_c_dbl_p)


class MatrixSingularException(Exception):
    pass


class MatrixNotSPDException(Exception):
    pass


def solve_symm_packed(ap: np.ndarray, b: np.ndarray, n: int, invert: bool) -> None:
    """MathExtension.solve(UpperSymmPackMatrix N, DenseVector n, numRows, invert), MathExtension.java:338-366.
    In place: ``b <- x`` and, if invert, ``ap <- N^-1`` (packed upper)."""
    assert ap.dtype == np.float64 and b.dtype == np.float64 and ap.flags.c_contiguous
    nn, nrhs, ldb, info = ctypes.c_int(n), ctypes.c_int(1), ctypes.c_int(max(1, n)), ctypes.c_int(0)
    ipiv = np.zeros(max(1, n), np.int32)
    ipp = ipiv.ctypes.data_as(_c_int_p)
    _dspsv(b'U', ctypes.byref(nn), ctypes.byref(nrhs), _p(ap), ipp, _p(b), ctypes.byref(ldb), ctypes.byref(info))
    if info.value > 0:
        raise MatrixSingularException()
    if info.value < 0:
        raise ValueError('dspsv illegal argument %d' % info.value)
    if invert:
        work = np.zeros(max(1, n))
        _dsptri(b'U', ctypes.byref(nn), _p(ap), ipp, _p(work), ctypes.byref(info))
        if info.value > 0:
            raise MatrixSingularException()
        if info.value < 0:
            raise ValueError('dsptri illegal argument %d' % info.value)


def inv_spd_packed(ap: np.ndarray, n: int) -> None:
    """MathExtension.inv(UpperSPDPackMatrix), MathExtension.java:304-324: dpptrf + dpptri in place."""
    nn, info = ctypes.c_int(n), ctypes.c_int(0)
    _dpptrf(b'U', ctypes.byref(nn), _p(ap), ctypes.byref(info))
    if info.value > 0:
        raise MatrixNotSPDException()
    if info.value < 0:
        raise ValueError('dpptrf illegal argument')
    _dpptri(b'U', ctypes.byref(nn), _p(ap), ctypes.byref(info))
    if info.value > 0:
        raise MatrixNotSPDException()
    if info.value < 0:
        raise ValueError('dpptri illegal argument')


def inv_symm_packed(ap: np.ndarray, n: int) -> None:
    """MathExtension.inv(UpperSymmPackMatrix), MathExtension.java:403-426: dsptrf + dsptri in place."""
    nn, info = ctypes.c_int(n), ctypes.c_int(0)
    ipiv = np.zeros(max(1, n), np.int32)
    ipp = ipiv.ctypes.data_as(_c_int_p)
    _dsptrf(b'U', ctypes.byref(nn), _p(ap), ipp, ctypes.byref(info))
    if info.value > 0:
        raise MatrixSingularException()
    if info.value < 0:
        raise ValueError('dsptrf illegal argument')
    work = np.zeros(max(1, n))
    _dsptri(b'U', ctypes.byref(nn), _p(ap), ipp, _p(work), ctypes.byref(info))
    if info.value > 0:
        raise MatrixSingularException()
    if info.value < 0:
        raise ValueError('dsptri illegal argument')
