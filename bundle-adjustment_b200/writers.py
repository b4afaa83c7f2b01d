"""Result writers on the new buffers.

Mirrors of the reference's ``AdjustmentResultWritable`` implementations
(``util/io/writer/MatlabResultWriter.java:52-226``, ``util/io/writer/DefaultResultWriter.java:46-155``): same file
names, variable names, index conventions and number formats.  The difference is where the numbers come from: only the
sub-matrix of Qxx that a writer exports is gathered -- on the device, by ``jaicov_get_qxx_submatrix`` -- so the full
cofactor matrix (16 GB packed at config 5, not representable in MTJ beyond n = 46 340) never travels to the host.
"""
from __future__ import annotations

import decimal

import numpy as np

from .host import COL_FIXED, PolynomialCoefficient


_CTX = decimal.Context(prec=800, rounding=decimal.ROUND_HALF_UP)


def java_format_f(value, width, precision, plus=False):
    """``String.format(Locale.ENGLISH, "%[+]<width>.<precision>f", double)`` as java.util.Formatter prints it (the reference needs
    JDK 25): the digits are those of ``Double.toString`` -- the SHORTEST decimal that round-trips -- rounded HALF_UP to the precision
    and padded with zeros beyond them.  C's printf (and Python's ``%``) expand the exact binary value instead and round it half-even:
    ``%.20f`` of 0.1 is ``0.10000000000000000000`` in Java but ``0.10000000000000000555`` in C, ``%.1f`` of 0.15 is ``0.2`` vs ``0.1``.
    The .info / .cxx files of DefaultResultWriter (formats ``%35.15f`` and ``%+35.15f``, DefaultResultWriter.java:67,142) carry up
    to 18 significant digits, so the difference shows in their last digits."""
    v = float(value)
    if v != v:
        s = 'NaN'
    elif v in (float('inf'), float('-inf')):
        s = ('-' if v < 0 else ('+' if plus else '')) + 'Infinity'
    else:
        q = _CTX.quantize(decimal.Decimal(repr(v)), decimal.Decimal(1).scaleb(-precision))
        s = format(q, 'f')
        if plus and not s.startswith('-'):
            s = '+' + s
    return s.rjust(width)


class BundleAdjustmentResultWriter:
    def __init__(self, exportPathAndFileBaseName):
        self._base = exportPathAndFileBaseName

    def getExportPathAndFileBaseName(self):
        return self._base

    def export(self, bundleAdjustment):
        raise NotImplementedError

    @staticmethod
    def _has_cofactor(adj):
        # exportDispersionMatrix = !(cofactor == null || cofactor.numRows() < u + d)
        return adj.getInvertNormalEquation().name != 'NONE' and adj._session is not None


def _point_indices(adj, first_index):
    """Columns of the object coordinates in getObjectCoordinates() order and their running export index
    (MatlabResultWriter.java:96-140 counts from 1, DefaultResultWriter.java:75-118 from 0)."""
    indices, cov = [], []
    k = first_index
    for oc in adj.getObjectCoordinates():
        row = []
        for p in (oc.getX(), oc.getY(), oc.getZ()):
            c = p.getColumn()
            if 0 <= c < COL_FIXED:
                indices.append(c)
                row.append(k)
                k += 1
            else:
                row.append(-1)
        cov.append(row)
    return indices, cov, k


class MatlabResultWriter(BundleAdjustmentResultWriter):
    """Writes ``<base>.mat`` (MAT5) with the reference's variables: variance_of_unit_weight_prio/post, degree_of_freedom,
    number_of_observations, number_of_unknowns, coordinates, interior_orientations, distortion_parameters, dispersion."""

    def export(self, bundleAdjustment):
        from scipy.io import savemat
        adj = bundleAdjustment
        if adj is None:
            raise ValueError('Error, bundle adjustment object cannot be null!')
        if self._base is None:
            raise ValueError('Error, export path cannot be null!')
        export_disp = self._has_cofactor(adj)
        n_cols = adj.getNumberOfUnknownParameters() + adj.getNumberOfDatumConditions()
        indices, cov, k = _point_indices(adj, 1)
        coords = np.zeros(len(cov), dtype=[('name', 'O'), ('X', 'f8'), ('Y', 'f8'), ('Z', 'f8'), ('covx', 'i4'), ('covy', 'i4'), ('covz', 'i4')])
        for i, (oc, c) in enumerate(zip(adj.getObjectCoordinates(), cov)):
            coords[i] = (oc.getName(), oc.getX().getValue(), oc.getY().getValue(), oc.getZ().getValue(), c[0], c[1], c[2])
        io_rows, dist_rows = [], []
        for cam in adj.getCameras():                                  # :143-160
            for p in cam.getInteriorOrientation():
                col = p.getColumn()
                if export_disp and 0 <= col < n_cols:
                    indices.append(col); c = k; k += 1
                else:
                    c = -1
                io_rows.append((cam.getId(), p.getParameterType().name.lower(), p.getValue(), c))
        for cam in adj.getCameras():                                  # :163-190
            for model in cam.getDistortionModels():
                for p in model:
                    order = p.getOrder() if isinstance(p, PolynomialCoefficient) else -1
                    col = p.getColumn()
                    if export_disp and 0 <= col < n_cols:
                        indices.append(col); c = k; k += 1
                    else:
                        c = -1
                    dist_rows.append((cam.getId(), p.getParameterType().name.lower(), p.getValue(), order, c))
        # the reference sets the "cov" field of these two structs only when a dispersion matrix is exported (:150-160, :175-187)
        cov_field = [('cov', 'i4')] if export_disp else []
        if not export_disp:
            io_rows, dist_rows = [r[:-1] for r in io_rows], [r[:-1] for r in dist_rows]
        io = np.array(io_rows, dtype=[('cam_id', 'i8'), ('name', 'O'), ('value', 'f8')] + cov_field)
        dist = np.array(dist_rows, dtype=[('cam_id', 'i8'), ('name', 'O'), ('value', 'f8'), ('order', 'i4')] + cov_field)
        out = {
            'variance_of_unit_weight_prio': adj.getVarianceFactorApriori(),
            'variance_of_unit_weight_post': adj.getVarianceFactorAposteriori(),
            'degree_of_freedom': np.int32(adj.getDegreeOfFreedom()),
            'number_of_observations': np.int32(adj.getNumberOfObservations()),
            'number_of_unknowns': np.int32(adj.getNumberOfUnknownParameters()),
            'coordinates': coords, 'interior_orientations': io, 'distortion_parameters': dist,
        }
        if export_disp:
            # the UNSCALED cofactor sub-matrix (:209-223)
            out['dispersion'] = adj._session.qxx_submatrix(np.array(indices, np.int32), 1.0)
        savemat(self._base + '.mat', out, format='5', oned_as='row')
        self.indices = indices
        return self._base + '.mat'


class DefaultResultWriter(BundleAdjustmentResultWriter):
    """Writes ``<base>.info`` (name, component, coordinate, row/column) and ``<base>.cxx`` (sigma0^2 * Qxx of the
    object coordinates), formats of DefaultResultWriter.java:67,142."""

    def export(self, bundleAdjustment):
        adj = bundleAdjustment
        if adj is None:
            raise ValueError('Error, bundle adjustment object cannot be null!')
        if self._base is None:
            raise ValueError('Error, export path cannot be null!')
        indices, cov, _ = _point_indices(adj, 0)
        with open(self._base + '.info', 'w') as f:
            for oc, c in zip(adj.getObjectCoordinates(), cov):
                for comp, p, ci in zip('XYZ', (oc.getX(), oc.getY(), oc.getZ()), c):
                    f.write('%25s\t%5s\t%s\t%10d\n' % (oc.getName(), comp, java_format_f(p.getValue(), 35, 15), ci))
        if self._has_cofactor(adj):
            C = adj._session.qxx_submatrix(np.array(indices, np.int32), adj.getVarianceFactorAposteriori())
            with open(self._base + '.cxx', 'w') as f:
                for row in C:
                    f.write(''.join(java_format_f(v, 35, 15, plus=True) + '  ' for v in row) + '\n')
        self.indices = indices
        return self._base + '.info', self._base + '.cxx'
