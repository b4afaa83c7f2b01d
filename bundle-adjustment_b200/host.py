"""Host-side mirror of the JAICOV API for the adjustment path (Python, because no JVM exists in this image).

Same names, argument meaning and error behaviour as the reference classes it stands in for
(paths relative to /root/reference/JAICOV/src/org/applied_geodesy/adjustment/):

  ParameterType                      bundle/parameter/ParameterType.java:27-110
  UnknownParameter / ObservationParameter   bundle/parameter/UnknownParameter.java:26-53, ObservationParameter.java:26-64
  ObjectCoordinate, ScaleBar         bundle/ObjectCoordinate.java:32-110, bundle/ScaleBar.java:30-39
  InteriorOrientation / ExteriorOrientation bundle/camera/orientation/*.java
  DistortionModel.Type and the model classes bundle/camera/distortion/*.java (default column = fixed for Bx,By,Cx,Cy)
  Camera, Image                      bundle/camera/Camera.java:38-139, bundle/camera/Image.java:32-90
  DirectlyObservedParameterGroup     bundle/parameter/DirectlyObservedParameterGroup.java:37-105
  BundleAdjustment                   bundle/BundleAdjustment.java (add :652-665, estimateModel :203, setters :1123-1195,
                                     getters :1048-1118, :1177)
  EstimationStateType, EstimationType, MatrixInversion

What happens here is what stays on the Java side in a real integration: the object graph, the
integer bookkeeping of prepareUnknownParameters (:667-782) and detectRankDefect (:836-1042), and flattening into the
C ABI.  All floating-point work of the adjustment runs in libjaicov_b200.so on the GPU.

Scale: object points and image observations may also be added in bulk (``ObjectCoordinateArray``, ``Image.addAll``)
so that 10^7 observations do not need 10^7 Python objects; the per-object calls of the reference API work too.
"""
from __future__ import annotations

import enum
import math

import numpy as np

from . import _lib

COL_UNSET, COL_FIXED = -1, 2147483647


class ParameterType(enum.IntEnum):
    PRINCIPAL_POINT_X = 111
    PRINCIPAL_POINT_Y = 112
    PRINCIPAL_DISTANCE = 113
    RADIAL_POLYNOMIAL_A = 121
    TANGENTIAL_POLYNOMIAL_B = 131
    TANGENTIAL_DISTORTION_Bx = 132
    TANGENTIAL_DISTORTION_By = 133
    AFFINITY_AND_SHEAR_Cx = 141
    AFFINITY_AND_SHEAR_Cy = 142
    DISTANCE_POLYNOMIAL_D = 151
    ZERNIKE_POLYNOMIAL_X = 161
    ZERNIKE_POLYNOMIAL_Y = 162
    ZERNIKE_POLYNOMIAL_Z = 163
    CAMERA_COORDINATE_X = 251
    CAMERA_COORDINATE_Y = 252
    CAMERA_COORDINATE_Z = 253
    CAMERA_OMEGA = 261
    CAMERA_PHI = 262
    CAMERA_KAPPA = 263
    OBJECT_COORDINATE_X = 311
    OBJECT_COORDINATE_Y = 312
    OBJECT_COORDINATE_Z = 313
    IMAGE_COORDINATE_X = 411
    IMAGE_COORDINATE_Y = 412
    SCALE_BAR_LENGTH = 511
    DIRECT_LINEAR_TRANSFORMATION_B11 = 611
    DIRECT_LINEAR_TRANSFORMATION_B12 = 612
    DIRECT_LINEAR_TRANSFORMATION_B13 = 613
    DIRECT_LINEAR_TRANSFORMATION_B14 = 614
    DIRECT_LINEAR_TRANSFORMATION_B21 = 621
    DIRECT_LINEAR_TRANSFORMATION_B22 = 622
    DIRECT_LINEAR_TRANSFORMATION_B23 = 623
    DIRECT_LINEAR_TRANSFORMATION_B24 = 624
    DIRECT_LINEAR_TRANSFORMATION_B31 = 631
    DIRECT_LINEAR_TRANSFORMATION_B32 = 632
    DIRECT_LINEAR_TRANSFORMATION_B33 = 633

    def getId(self):
        return int(self)


class EstimationStateType(enum.Enum):
    ERROR_FREE_ESTIMATION = 1
    BUSY = 0
    INTERRUPT = -1
    SINGULAR_MATRIX = -2
    ROBUST_ESTIMATION_FAILED = -3
    NO_CONVERGENCE = -4
    NOT_INITIALISED = -5
    EXPORT_ADJUSTMENT_RESULTS_FAILED = -6
    OUT_OF_MEMORY = -7

    def getId(self):
        return self.value


class EstimationType(enum.Enum):
    L1NORM = 1
    L2NORM = 2
    SIMULATION = 3
    MODIFIED_UNSCENTED_TRANSFORMATION = 4
    SPHERICAL_SIMPLEX_UNSCENTED_TRANSFORMATION = 5


class MatrixInversion(enum.Enum):
    NONE = 0
    FULL = 1
    PRE_ELIMINATION = 2
    REDUCED = 3


class UnknownParameter:
    """Value + column (-1 = not set, Integer.MAX_VALUE = fixed), parameter/UnknownParameter.java:26-53."""
    __slots__ = ('_type', '_ref', '_value', '_column')

    def __init__(self, parameterType, reference=None, value=0.0):
        self._type, self._ref, self._value, self._column = parameterType, reference, float(value), COL_UNSET

    def getParameterType(self): return self._type
    def getReference(self): return self._ref
    def getValue(self): return self._value
    def setValue(self, v): self._value = float(v)
    def getColumn(self): return self._column
    def setColumn(self, c): self._column = int(c)


class PolynomialCoefficient(UnknownParameter):
    __slots__ = ('_order',)

    def __init__(self, parameterType, reference, order):
        super().__init__(parameterType, reference)
        self._order = int(order)

    def getOrder(self): return self._order


class ObservationParameter:
    """parameter/ObservationParameter.java:26-64 (value, variance, row; reference = observed UnknownParameter)."""
    __slots__ = ('_type', '_ref', '_value', '_variance', '_row')

    def __init__(self, parameter, value=None, variance=None):
        self._ref = parameter
        self._type = parameter.getParameterType()
        self._value = parameter.getValue() if value is None else float(value)
        self._variance = None
        self._row = -1
        if variance is not None:
            self.setVariance(variance)

    def getParameterType(self): return self._type
    def getReference(self): return self._ref
    def getValue(self): return self._value
    def setValue(self, v): self._value = float(v)
    def getVariance(self): return self._variance
    def getRow(self): return self._row

    def setVariance(self, variance):
        if not variance > 0:
            raise ValueError('Error, variance must be positive: %r' % (variance,))
        self._variance = float(variance)


# ---- object points ----------------------------------------------------------------------------------------------------
class ObjectCoordinateArray:
    """Bulk storage of object points: what n ObjectCoordinate instances hold (bundle/ObjectCoordinate.java:32-47)."""

    def __init__(self, names, xyz):
        self.xyz = np.array(xyz, dtype=np.float64).reshape(-1, 3)
        n = self.xyz.shape[0]
        self.names = list(names) if names is not None else [str(i) for i in range(n)]
        self.column = np.full((n, 3), COL_UNSET, dtype=np.int64)
        self.datum = np.ones(n, dtype=bool)          # datum = TRUE by default (:34)

    def __len__(self): return self.xyz.shape[0]
    def __getitem__(self, i): return ObjectCoordinate._view(self, int(i))


class _PointParameter:
    """UnknownParameter view of one coordinate component of an ObjectCoordinateArray."""
    __slots__ = ('_store', '_i', '_c')

    def __init__(self, store, i, c): self._store, self._i, self._c = store, i, c
    def getParameterType(self): return (ParameterType.OBJECT_COORDINATE_X, ParameterType.OBJECT_COORDINATE_Y, ParameterType.OBJECT_COORDINATE_Z)[self._c]
    def getReference(self): return ObjectCoordinate._view(self._store, self._i)
    def getValue(self): return float(self._store.xyz[self._i, self._c])
    def setValue(self, v): self._store.xyz[self._i, self._c] = v
    def getColumn(self): return int(self._store.column[self._i, self._c])
    def setColumn(self, c): self._store.column[self._i, self._c] = c


class ObjectCoordinate:
    __slots__ = ('_store', '_i')

    def __init__(self, name, x, y, z):
        self._store = ObjectCoordinateArray([name], [[x, y, z]])
        self._i = 0

    @classmethod
    def _view(cls, store, i):
        o = cls.__new__(cls)
        o._store, o._i = store, i
        return o

    def getName(self): return self._store.names[self._i]
    def getX(self): return _PointParameter(self._store, self._i, 0)
    def getY(self): return _PointParameter(self._store, self._i, 1)
    def getZ(self): return _PointParameter(self._store, self._i, 2)
    def isDatum(self): return bool(self._store.datum[self._i])
    def setDatum(self, datum): self._store.datum[self._i] = bool(datum)
    def __eq__(self, o): return isinstance(o, ObjectCoordinate) and o._store is self._store and o._i == self._i
    def __hash__(self): return hash((id(self._store), self._i))


class ScaleBar:
    """bundle/ScaleBar.java:34-39: variance = sigma^2."""

    def __init__(self, objectCoordinateA, objectCoordinateB, value, sigma):
        self._a, self._b = objectCoordinateA, objectCoordinateB
        self._length = ObservationParameter(UnknownParameter(ParameterType.SCALE_BAR_LENGTH, self), value, sigma * sigma)

    def getLength(self): return self._length
    def getObjectCoordinateA(self): return self._a
    def getObjectCoordinateB(self): return self._b


# ---- camera ---------------------------------------------------------------------------------------------------------------
class InteriorOrientation:
    def __init__(self, camera):
        self._camera = camera
        self._x0 = UnknownParameter(ParameterType.PRINCIPAL_POINT_X, self)
        self._y0 = UnknownParameter(ParameterType.PRINCIPAL_POINT_Y, self)
        self._c = UnknownParameter(ParameterType.PRINCIPAL_DISTANCE, self)

    def getPrinciplePointX(self): return self._x0
    def getPrinciplePointY(self): return self._y0
    def getPrincipleDistance(self): return self._c
    def getReference(self): return self._camera
    def __iter__(self): return iter((self._x0, self._y0, self._c))   # iterator order :70-79


class ExteriorOrientation:
    _ORDER = (ParameterType.CAMERA_COORDINATE_X, ParameterType.CAMERA_COORDINATE_Y, ParameterType.CAMERA_COORDINATE_Z,
              ParameterType.CAMERA_OMEGA, ParameterType.CAMERA_PHI, ParameterType.CAMERA_KAPPA)

    def __init__(self, image):
        self._image = image
        self._params = {t: UnknownParameter(t, self) for t in self._ORDER}

    def get(self, parameterType): return self._params[parameterType]
    def getReference(self): return self._image
    def __iter__(self): return iter(self._params.values())


class DistortionModel:
    class Type(enum.IntEnum):   # enum ordinal order, camera/distortion/DistortionModel.java:29-37
        AFFINITY_AND_SHEAR = 0
        TANGENTIAL_DISTORTION = 1
        RADIAL_DISTORTION = 2
        DISTANCE_DISTORTION = 3
        ZERNIKE_X = 4
        ZERNIKE_Y = 5
        ZERNIKE_GRADIENT = 6

    def __init__(self, camera, type_, r0=0.0):
        self._camera, self._type, self._r0 = camera, type_, float(r0)
        self._params = {}    # insertion ordered: order -> parameter

    def getType(self): return self._type
    def getReference(self): return self._camera
    def getR0(self): return self._r0
    def getNumberOfParameters(self): return len(self._params)
    def __iter__(self): return iter(self._params.values())

    def _add(self, order, coefficient):
        if order in self._params:
            raise ValueError('Error, polynomial coefficient order already exists. %d' % order)
        self._params[order] = coefficient
        return coefficient

    def get(self, order): return self._params.get(order)


class AffinityShearDistortionModel(DistortionModel):
    def __init__(self, camera):
        super().__init__(camera, DistortionModel.Type.AFFINITY_AND_SHEAR)
        self._cx = UnknownParameter(ParameterType.AFFINITY_AND_SHEAR_Cx, self)
        self._cy = UnknownParameter(ParameterType.AFFINITY_AND_SHEAR_Cy, self)
        self._cx.setColumn(COL_FIXED)      # camera/distortion/AffinityShearDistortionModel.java:39-40
        self._cy.setColumn(COL_FIXED)

    def getCx(self): return self._cx
    def getCy(self): return self._cy
    def getNumberOfParameters(self): return 2
    def __iter__(self): return iter((self._cx, self._cy))


class TangentialDistortionModel(DistortionModel):
    def __init__(self, camera):
        super().__init__(camera, DistortionModel.Type.TANGENTIAL_DISTORTION)
        self._bx = UnknownParameter(ParameterType.TANGENTIAL_DISTORTION_Bx, self)
        self._by = UnknownParameter(ParameterType.TANGENTIAL_DISTORTION_By, self)
        self._bx.setColumn(COL_FIXED)      # camera/distortion/TangentialDistortionModel.java:38-39
        self._by.setColumn(COL_FIXED)
        self._add(-1, self._bx)
        self._add(-2, self._by)

    def getBx(self): return self._bx
    def getBy(self): return self._by

    def add(self, order):
        if order <= 0:
            raise ValueError('Error, polynomial coefficient order must be a real positive integer. %d' % order)
        return self._add(order, PolynomialCoefficient(ParameterType.TANGENTIAL_POLYNOMIAL_B, self, order))


class RadiallySymmetricDistortionModel(DistortionModel):
    def __init__(self, camera, r0):
        super().__init__(camera, DistortionModel.Type.RADIAL_DISTORTION, r0)

    def add(self, order):
        if order <= 0:
            raise ValueError('Error, polynomial coefficient order must be a real positive integer. %d' % order)
        return self._add(order, PolynomialCoefficient(ParameterType.RADIAL_POLYNOMIAL_A, self, order))


class RadialDistanceDistortionModel(DistortionModel):
    def __init__(self, camera, r0):
        super().__init__(camera, DistortionModel.Type.DISTANCE_DISTORTION, r0)

    def add(self, order):
        if order <= 0:
            raise ValueError('Error, polynomial coefficient order must be a real positive integer. %d' % order)
        return self._add(order, PolynomialCoefficient(ParameterType.DISTANCE_POLYNOMIAL_D, self, order))


class ZernikeDistortionModel(DistortionModel):
    _PT = {DistortionModel.Type.ZERNIKE_X: ParameterType.ZERNIKE_POLYNOMIAL_X,
           DistortionModel.Type.ZERNIKE_Y: ParameterType.ZERNIKE_POLYNOMIAL_Y,
           DistortionModel.Type.ZERNIKE_GRADIENT: ParameterType.ZERNIKE_POLYNOMIAL_Z}

    def add(self, order):
        if order <= 0:
            raise ValueError('Error, polynomial coefficient order must be a real positive integer. %d' % order)
        return self._add(order, PolynomialCoefficient(self._PT[self._type], self, order))


class Image:
    """camera/Image.java:32-90.  Observations are kept as array chunks (store, indices, xy, sigma, rho)."""

    def __init__(self, id_, camera):
        self._id, self._camera = id_, camera
        self._eo = ExteriorOrientation(self)
        self._chunks = []
        self._seen = set()
        self._dispersion = None

    def setDispersion(self, packed_upper):
        """EXTENSION (not in the reference, whose image-coordinate groups are the two rows of one point, camera/ImageCoordinate.java:
        102-104): fully populated dispersion of the image's 2m coordinates (x_0, y_0, x_1, y_1, ... in observation order), MTJ packed
        upper; replaces the per-point sigma / rho.  None removes it."""
        self._dispersion = None if packed_upper is None else np.ascontiguousarray(packed_upper, np.float64)

    def getDispersion(self): return self._dispersion

    def getId(self): return self._id
    def getReference(self): return self._camera
    def getExteriorOrientation(self): return self._eo
    def getNumberOfImageCoordinates(self): return sum(len(c[1]) for c in self._chunks)

    def add(self, objectCoordinate, xp, yp, sigmax, sigmay, corrCoefXY=0.0):
        if abs(corrCoefXY) >= 1:   # camera/ImageCoordinate.java:41-42
            raise ValueError('Error, correlation coefficient rho(x,y) must be in the open interval (-1 1): %r' % corrCoefXY)
        if sigmax * sigmax <= 0 or sigmay * sigmay <= 0:
            raise ValueError('Error, variance must be positive')
        key = (id(objectCoordinate._store), objectCoordinate._i)
        if key in self._seen:      # camera/Image.java:56-58: a second observation of the same point is ignored
            return
        self._seen.add(key)
        if self._chunks and self._chunks[-1][0] is objectCoordinate._store and isinstance(self._chunks[-1][1], list):
            ch = self._chunks[-1]
            ch[1].append(objectCoordinate._i); ch[2].append((xp, yp)); ch[3].append((sigmax, sigmay)); ch[4].append(corrCoefXY)
        else:
            self._chunks.append([objectCoordinate._store, [objectCoordinate._i], [(xp, yp)], [(sigmax, sigmay)], [corrCoefXY]])

    def get(self, objectCoordinate):
        """camera/Image.java:77-79: the image coordinate of an object point, or None if the point is not observed here."""
        if getattr(self, '_lookup_n', -1) != len(self._chunks):
            self._lookup = {}
            for (store, idx, xy_c, _sg, _rh) in self._chunks:
                for k, i in enumerate(np.asarray(idx, np.int64).tolist()):
                    self._lookup.setdefault((id(store), i), tuple(np.asarray(xy_c, float).reshape(-1, 2)[k]))
            self._lookup_n = len(self._chunks)
            if self._chunks and isinstance(self._chunks[-1][1], list):
                self._lookup_n = -1            # an open add() chunk may still grow: rebuild next time
        return self._lookup.get((id(objectCoordinate._store), objectCoordinate._i))

    def addAll(self, objectCoordinates, indices, xy, sigma, rho=None):
        """Bulk form of add(): points ``objectCoordinates[indices[k]]`` (distinct) observed at xy[k] with sigma[k], rho[k]."""
        indices = np.asarray(indices, np.int64)
        xy = np.asarray(xy, float).reshape(-1, 2)
        sigma = np.broadcast_to(np.asarray(sigma, float), xy.shape)
        rho = np.zeros(len(indices)) if rho is None else np.asarray(rho, float)
        if np.any(np.abs(rho) >= 1):
            raise ValueError('Error, correlation coefficient rho(x,y) must be in the open interval (-1 1)')
        if np.any(sigma <= 0):
            raise ValueError('Error, variance must be positive')
        self._chunks.append([objectCoordinates, indices, xy, sigma, rho])


class Camera:
    """camera/Camera.java:38-139; models are kept in enum-ordinal order (Arrays.sort, :50)."""

    def __init__(self, id_, r0, *distortionModelTypes):
        self._id = id_
        self._io = InteriorOrientation(self)
        self._images = {}
        self._models = {}
        types = sorted(distortionModelTypes)
        for t in types:
            if t in self._models:
                raise ValueError('Error, duplicate type of distortion model detected. %s' % t)
            T = DistortionModel.Type
            if t == T.AFFINITY_AND_SHEAR: m = AffinityShearDistortionModel(self)
            elif t == T.TANGENTIAL_DISTORTION: m = TangentialDistortionModel(self)
            elif t == T.RADIAL_DISTORTION: m = RadiallySymmetricDistortionModel(self, r0)
            elif t == T.DISTANCE_DISTORTION: m = RadialDistanceDistortionModel(self, r0)
            else: m = ZernikeDistortionModel(self, t, r0)
            self._models[t] = m
        self._r0 = float(r0)

    def getId(self): return self._id
    def getInteriorOrientation(self): return self._io
    def getNumberOfImages(self): return len(self._images)
    def getDistortionModel(self, type_): return self._models.get(type_)
    def getDistortionModels(self): return list(self._models.values())
    def __iter__(self): return iter(self._images.values())

    def add(self, imageId):
        if imageId not in self._images:
            self._images[imageId] = Image(imageId, self)
        return self._images[imageId]


class DirectlyObservedParameterGroup:
    """parameter/DirectlyObservedParameterGroup.java:37-105.  ``dispersionMatrix``: None (diagonal, variances of
    the ObservationParameters), a packed-upper vector of length r(r+1)/2 (MTJ UpperSPDPackMatrix data) or a dense
    symmetric r x r array."""

    def __init__(self, observedParameters, dispersionMatrix=None):
        self._obs = list(observedParameters)
        if len(set(map(id, self._obs))) != len(self._obs):
            raise ValueError('Error, array contains duplicate observations.')
        self._packed = None
        if dispersionMatrix is not None:
            D = np.asarray(dispersionMatrix, float)
            r = len(self._obs)
            if D.ndim == 2:
                if D.shape != (r, r):
                    raise ValueError('Error, number of observations and number of rows/columns in dispersion matrix are unequal: %d vs. %d' % (r, D.shape[1]))
                iu = np.triu_indices(r)
                packed = np.empty(r * (r + 1) // 2)
                packed[iu[0] + iu[1] * (iu[1] + 1) // 2] = D[iu]
            else:
                if D.size != r * (r + 1) // 2:
                    raise ValueError('Error, number of observations and number of rows/columns in dispersion matrix are unequal')
                packed = D.copy()
            idx = np.arange(r)
            for o, v in zip(self._obs, packed[idx + idx * (idx + 1) // 2]):
                o.setVariance(float(v))
            self._packed = packed

    def hasFullyPopulatedWeightMatrix(self): return self._packed is not None
    def getNumberOfParameters(self): return len(self._obs)
    def __iter__(self): return iter(self._obs)


class UpperSymmPackMatrix:
    """Read-only stand-in for no.uib.cipr.matrix.UpperSymmPackMatrix holding Qxx: column-major packed upper,
    element (r,c), r<=c, at r + c(c+1)/2."""

    def __init__(self, n, data, adjustment=None):
        self._n, self._data = n, data
        self._adjustment = adjustment      # the BundleAdjustment whose device-resident Qxx this is a host copy of

    def numRows(self): return self._n
    def numColumns(self): return self._n
    def getData(self): return self._data

    def get(self, r, c):
        if r > c:
            r, c = c, r
        return float(self._data[r + c * (c + 1) // 2])

    def toDense(self):
        n = self._n
        D = np.empty((n, n))
        iu = np.triu_indices(n)
        D[iu] = self._data[iu[0] + iu[1] * (iu[1] + 1) // 2]
        D.T[iu] = D[iu]
        return D


class DLTCoefficients:
    """dlt/DLTCoefficients.java:33-84: the 11 DLT coefficients of an image plus the interior / exterior orientation derived
    from them, in the reference's insertion order."""
    _ORDER = tuple(ParameterType(i) for i in (611, 612, 613, 614, 621, 622, 623, 624, 631, 632, 633)) + (
        ParameterType.PRINCIPAL_POINT_X, ParameterType.PRINCIPAL_POINT_Y, ParameterType.PRINCIPAL_DISTANCE,
        ParameterType.CAMERA_COORDINATE_X, ParameterType.CAMERA_COORDINATE_Y, ParameterType.CAMERA_COORDINATE_Z,
        ParameterType.CAMERA_OMEGA, ParameterType.CAMERA_PHI, ParameterType.CAMERA_KAPPA)

    def __init__(self, image):
        self._image = image
        self._params = {t: UnknownParameter(t, self) for t in self._ORDER}

    def get(self, parameterType): return self._params[parameterType]
    def getReference(self): return self._image
    def __iter__(self): return iter(self._params.values())


class DirectLinearTransformation:
    """dlt/DirectLinearTransformation.java:49-184.  ``adjust`` keeps the reference's per-image signature; ``adjustAll`` is the
    batched form (one kernel launch for all images, jaicov_dlt_batch)."""

    class RestrictionType(enum.IntEnum):      # ordinal order, :50-57
        IDENTICAL_PRINCIPLE_DISTANCE = 0
        ROTATION_WITHOUT_SHEAR = 1
        FIXED_PRINCIPLE_DISTANCE_X = 2
        FIXED_PRINCIPLE_DISTANCE_Y = 3
        FIXED_PRINCIPAL_POINT_X = 4
        FIXED_PRINCIPAL_POINT_Y = 5

    maximalNumberOfIterations = 5000          # DefaultValue.getMaximalNumberOfIterations(), :63

    @staticmethod
    def adjust(coefficients, objectCoordinates, *restrictions, device=0):
        return DirectLinearTransformation.adjustAll([coefficients], objectCoordinates, *restrictions, device=device)[0]

    @staticmethod
    def adjustAll(coefficientsList, objectCoordinates, *restrictions, device=0):
        """objectCoordinates: {name: ObjectCoordinate} of the points with known coordinates.  Returns one bool per image."""
        PT = ParameterType
        pt_ptr, xy, xyz, io = [0], [], [], []
        for coef in coefficientsList:
            # prepareUnknwonParameters, :279-314
            column = 0
            for p in coef:
                p.setValue(0)
                if 611 <= int(p.getParameterType()) <= 633:
                    p.setColumn(column)
                    column += 1
                else:
                    p.setColumn(COL_FIXED if p.getColumn() == COL_FIXED else COL_UNSET)
            image = coef.getReference()
            inner = image.getReference().getInteriorOrientation()
            for p in (inner.getPrincipleDistance(), inner.getPrinciplePointX(), inner.getPrinciplePointY()):
                coef.get(p.getParameterType()).setValue(p.getValue())
                if p.getColumn() == COL_FIXED:
                    coef.get(p.getParameterType()).setColumn(COL_FIXED)
            # homologous points by name, :78-94
            cnt = 0
            for (store, idx, xy_c, _sg, _rh) in image._chunks:
                xy_c = np.asarray(xy_c, float).reshape(-1, 2)
                for k, i in enumerate(np.asarray(idx, np.int64).tolist()):
                    oc = objectCoordinates.get(store.names[i])
                    if oc is None:
                        continue
                    xy.append(xy_c[k])
                    xyz.append((oc.getX().getValue(), oc.getY().getValue(), oc.getZ().getValue()))
                    cnt += 1
            pt_ptr.append(pt_ptr[-1] + cnt)
            io.append((coef.get(PT.PRINCIPAL_DISTANCE).getValue(), coef.get(PT.PRINCIPAL_POINT_X).getValue(),
                       coef.get(PT.PRINCIPAL_POINT_Y).getValue()))
        out, status, _passes = _lib.dlt_batch(pt_ptr, np.array(xy).reshape(-1, 2), np.array(xyz).reshape(-1, 3), np.array(io).reshape(-1, 3),
                                              [int(r) for r in restrictions], DirectLinearTransformation.maximalNumberOfIterations, device)
        results = []
        for coef, o, st in zip(coefficientsList, out, status):
            if st < 0:
                results.append(False)
                continue
            for t, v in zip(DLTCoefficients._ORDER[:11], o[:11]):
                coef.get(t).setValue(v)
            # :248-265: fixed interior orientation values are kept
            for t, v in ((PT.PRINCIPAL_DISTANCE, o[11]), (PT.PRINCIPAL_POINT_X, o[12]), (PT.PRINCIPAL_POINT_Y, o[13])):
                if coef.get(t).getColumn() != COL_FIXED:
                    coef.get(t).setValue(v)
            for t, v in zip(DLTCoefficients._ORDER[14:], o[14:20]):
                coef.get(t).setValue(v)
            results.append(bool(st == 1))
        return results


class CoordinateTransformationExteriorOrientation:
    """tranformation/CoordinateTransformationExteriorOrientation.java: object points carried into the frame of a reference
    image, X_trg = X0_trg + R_trg R_src' (X - X0_src), with the propagated covariance sigma2 * J Qxx J' (:49-121).
    The contraction with Qxx runs on the device that holds it (jaicov_propagate_eo_transform); CoVar must be the matrix
    handed out by BundleAdjustment.getCofactorMatrix()."""
    _instance = None

    @classmethod
    def getInstance(cls):
        if cls._instance is None:
            cls._instance = cls()
        return cls._instance

    def __init__(self):
        self._covariance, self._transformed = None, []

    def transform(self, objectCoordinatesToTransform, imagesToAlign, sigma2, CoVar):
        adj = getattr(CoVar, '_adjustment', None)
        if adj is None or adj._session is None:
            raise _lib.JaicovError(_lib.NOT_INITIALISED, 'CoVar is not the cofactor matrix of a finished adjustment on the device')
        points, src, trg, names = [], [], [], []
        for referenceImage, images in imagesToAlign.items():              # :81-105
            for image in images:
                for oc in objectCoordinatesToTransform:
                    if image.get(oc) is None:                            # skip points not visible in the current image, :91-95
                        continue
                    points.append(adj._store_base[id(oc._store)] + oc._i)
                    src.append(adj._image_index[id(image)])
                    trg.append(adj._image_index[id(referenceImage)])
                    names.append('%s %s %s' % (oc.getName(), image.getId(), referenceImage.getId()))   # :103
        xyz, cov = adj._session.propagate_eo_transform(points, src, trg, sigma2)
        self._transformed = []
        for k, name in enumerate(names):
            t = ObjectCoordinate(name, *xyz[k])
            t.getX().setColumn(3 * k); t.getY().setColumn(3 * k + 1); t.getZ().setColumn(3 * k + 2)   # :146-148, :217-219
            self._transformed.append(t)
        self._covariance = UpperSymmPackMatrix(3 * len(names), cov)
        self._triples = (points, src, trg)

    def getCovarianceMatrix(self): return self._covariance
    def getTransformedCoordinates(self): return self._transformed


# ---- the adjustment ---------------------------------------------------------------------------------------------------
class BundleAdjustment:
    MatrixInversion = MatrixInversion

    def __init__(self, device=0):
        self._cameras, self._scaleBars, self._groups = [], [], []
        self._estimationType = EstimationType.L2NORM
        self._invert = MatrixInversion.FULL          # :91
        self._maxIter = 5000                          # DefaultValue.java:25
        self._applyAposteriori = True                 # :86
        self._useCentroid = True                      # :87
        self._damping = 0.0
        self._device = device
        self._listeners = []
        self._session = None
        self._Qxx = None
        self._numObs = self._numUnknown = self._defect = 0
        self._sigma2apriori = 1.0                     # :98
        self._omega = 0.0
        self._objectCoordinates = []
        self._status = EstimationStateType.BUSY
        self._writer = None
        self.stats = None

    # :652-665
    def add(self, *items):
        for it in items:
            if isinstance(it, Camera): self._cameras.append(it)
            elif isinstance(it, ScaleBar):
                if it not in self._scaleBars: self._scaleBars.append(it)
            elif isinstance(it, DirectlyObservedParameterGroup):
                if it not in self._groups: self._groups.append(it)
            else: raise TypeError('unsupported argument %r' % (it,))

    def addPropertyChangeListener(self, listener): self._listeners.append(listener)          # :1459-1461

    def removePropertyChangeListener(self, listener):                                         # :1463-1465
        if listener in self._listeners:
            self._listeners.remove(listener)

    def interrupt(self):                                                                      # :1455-1457
        """Stops a running estimateModel() at its next check (:240, :320); a request made while no adjustment runs
        takes effect in the next one, as in the reference."""
        self._interrupt_requested = True
        if self._session is not None:
            self._session.interrupt()

    def setAdjustmentResultWriter(self, adjustmentResultWriter):     # :1123-1125
        self._writer = adjustmentResultWriter

    def setEstimationType(self, estimationType):       # :1132-1137
        if estimationType in (EstimationType.L2NORM, EstimationType.SIMULATION):
            self._estimationType = estimationType
        else:
            raise NotImplementedError('BundleAdjustment Error, this estimation type is not supported! %s' % estimationType)

    def setInvertNormalEquation(self, invert): self._invert = invert
    def getInvertNormalEquation(self): return self._invert
    def useCentroidedCoordinates(self, flag): self._useCentroid = bool(flag)
    def applyAposterioriVarianceOfUnitWeight(self, flag): self._applyAposteriori = bool(flag)
    def setLevenbergMarquardtDampingValue(self, lam): self._damping = abs(lam)
    def getLevenbergMarquardtDampingValue(self): return self._damping
    def setMaximalNumberOfIterations(self, n): self._maxIter = int(n)

    def setSolver(self, solver):
        """Not part of the reference API: _lib.SOLVER_AUTO (default) / SOLVER_DENSE / SOLVER_STRUCTURED, see
        include/jaicov_b200.h.  Both routes return the reference's results."""
        self._solver = int(solver)

    def getNumberOfObservations(self): return self._numObs
    def getNumberOfUnknownParameters(self): return self._numUnknown
    def getNumberOfDatumConditions(self): return self._defect
    def getDegreeOfFreedom(self): return self._numObs - self._numUnknown + self._defect   # :1080-1082
    def getVarianceFactorApriori(self): return self._sigma2apriori
    def getCameras(self): return self._cameras
    def getScaleBars(self): return self._scaleBars
    def getObjectCoordinates(self): return self._objectCoordinates

    def getVarianceFactorAposteriori(self):             # :1090-1093
        dof = self.getDegreeOfFreedom()
        if dof > 0 and self._omega > 0 and self._estimationType != EstimationType.SIMULATION and self._applyAposteriori:
            return abs(self._omega / float(dof))
        return self._sigma2apriori

    def getCofactorMatrix(self):                        # :1177-1179
        if self._invert == MatrixInversion.NONE or self._session is None:
            return None
        if self._Qxx is None:
            n, nq = self._session.n, self._session.n_qxx
            data = self._session.qxx_packed()
            if nq < n:
                # REDUCED / PRE_ELIMINATION: the reference hands out the full-size matrix object whose leading block is
                # the reduced cofactor matrix (the rest holds unspecified leftovers there, zeros here)
                full = np.zeros(n * (n + 1) // 2)
                full[:data.size] = data
                data = full
            self._Qxx = UpperSymmPackMatrix(n, data, self)
        return self._Qxx

    # ---- prepareUnknownParameters, :667-782, and detectRankDefect, :836-1042 ---------------------------------------
    def _prepare(self):
        stores, base = [], {}

        def gidx(store, idx):
            if id(store) not in base:
                base[id(store)] = sum(len(s) for s in stores)
                stores.append(store)
            return base[id(store)] + np.asarray(idx, np.int64)

        sigma2 = self._sigma2apriori
        cam_of_img, eo_params, pt_ptr = [], [], [0]
        obj, xy, var, rho = [], [], [], []
        images = []
        img_sigma = []
        for ci, cam in enumerate(self._cameras):
            for img in cam:
                if img.getDispersion() is not None:
                    sg = img.getDispersion()
                    m2 = 2 * img.getNumberOfImageCoordinates()
                    if sg.size != m2 * (m2 + 1) // 2:
                        raise ValueError('Error, the dispersion of an image needs (2m)(2m+1)/2 entries')
                    k = np.arange(m2, dtype=np.int64)
                    sigma2 = min(sigma2, float(sg[k + k * (k + 1) // 2].min()))     # like DirectlyObservedParameterGroup, DOPG:55-57
                    img_sigma.append((len(images), sg))
                images.append(img)
                cam_of_img.append(ci)
                eo_params.append(list(img.getExteriorOrientation()))
                cnt = 0
                for (store, idx, xy_c, sg_c, rh_c) in img._chunks:
                    g = gidx(store, idx)
                    s = np.asarray(sg_c, float).reshape(-1, 2)
                    obj.append(g); xy.append(np.asarray(xy_c, float).reshape(-1, 2)); var.append(s * s)
                    rho.append(np.asarray(rh_c, float).reshape(-1))
                    cnt += g.size
                    if g.size:
                        sigma2 = min(sigma2, float((s * s).min()))
                pt_ptr.append(pt_ptr[-1] + cnt)
        obj = np.concatenate(obj) if obj else np.zeros(0, np.int64)
        numObs = 2 * obj.size
        for bar in self._scaleBars:
            for oc in (bar.getObjectCoordinateA(), bar.getObjectCoordinateB()):
                gidx(oc._store, [oc._i])
        for grp in self._groups:
            for op in grp:
                ref = op.getReference()
                if isinstance(ref, _PointParameter):
                    gidx(ref._store, [ref._i])
        nPt = sum(len(s) for s in stores)
        pt_col = np.concatenate([s.column for s in stores]) if stores else np.zeros((0, 3), np.int64)
        pt_col = pt_col.copy()
        datum = np.concatenate([s.datum for s in stores]) if stores else np.zeros(0, bool)
        counter = 0
        # object coordinates in order of first appearance (:677-683)
        if obj.size:
            _, first = np.unique(obj, return_index=True)
            oc_order = obj[np.sort(first)]
        else:
            oc_order = np.zeros(0, np.int64)
        sub = pt_col[oc_order]
        new = sub == COL_UNSET
        sub[new] = counter + np.arange(int(new.sum()))
        counter += int(new.sum())
        pt_col[oc_order] = sub
        in_adjustment = np.zeros(nPt, bool)
        in_adjustment[oc_order] = True
        oc_list = list(oc_order.tolist())
        # interior orientation and distortion parameters (:695-713)
        cam_params = []
        nIO = nDist = 0
        for cam in self._cameras:
            plist = list(cam.getInteriorOrientation())
            for p in plist:
                if p.getColumn() == COL_UNSET:
                    nIO += 1
            for model in cam.getDistortionModels():
                for p in model:
                    if p.getColumn() == COL_UNSET:
                        nDist += 1
                    plist.append(p)
            cam_params.append(plist)
        assigned = []   # (parameter, column) for UnknownParameter objects

        def add_unknown(p):
            nonlocal counter
            if p.getColumn() == COL_UNSET and id(p) not in seen:
                seen[id(p)] = counter
                assigned.append(p)
                counter += 1
        seen = {}
        for plist in cam_params:
            for p in plist:
                add_unknown(p)
        for ep in eo_params:                      # :715-722
            for p in ep:
                add_unknown(p)

        def add_point(g):
            nonlocal counter
            if not in_adjustment[g]:
                in_adjustment[g] = True
                oc_list.append(int(g))
            for c in range(3):
                if pt_col[g, c] == COL_UNSET:
                    pt_col[g, c] = counter
                    counter += 1
        bar_a, bar_b, bar_len, bar_var = [], [], [], []
        for bar in self._scaleBars:               # :724-745
            numObs += 1
            a = int(gidx(bar.getObjectCoordinateA()._store, [bar.getObjectCoordinateA()._i])[0])
            b = int(gidx(bar.getObjectCoordinateB()._store, [bar.getObjectCoordinateB()._i])[0])
            for g in (a, b):
                if not in_adjustment[g]:
                    in_adjustment[g] = True
                    oc_list.append(g)
            add_point(a); add_point(b)
            bar_a.append(a); bar_b.append(b); bar_len.append(bar.getLength().getValue()); bar_var.append(bar.getLength().getVariance())
            sigma2 = min(sigma2, bar.getLength().getVariance())
        param_slot = {}
        for ci, plist in enumerate(cam_params):
            for k, p in enumerate(plist):
                param_slot[id(p)] = (1, ci, k) if k < 3 else (2, ci, k - 3)
        for ii, ep in enumerate(eo_params):
            for k, p in enumerate(ep):
                param_slot[id(p)] = (3, ii, k)
        groups = []
        for grp in self._groups:                  # :747-771
            kind, index, comp, obs, gvar = [], [], [], [], []
            for op in grp:
                ref = op.getReference()
                if isinstance(ref, _PointParameter):
                    g = int(gidx(ref._store, [ref._i])[0])
                    if not in_adjustment[g]:
                        in_adjustment[g] = True
                        oc_list.append(g)
                    if pt_col[g, ref._c] == COL_UNSET:
                        pt_col[g, ref._c] = counter
                        counter += 1
                    kind.append(0); index.append(g); comp.append(ref._c)
                else:
                    if id(ref) not in param_slot:
                        raise ValueError('observed parameter does not belong to a camera or image of this adjustment')
                    add_unknown(ref)
                    k, ix, cp = param_slot[id(ref)]
                    kind.append(k); index.append(ix); comp.append(cp)
                obs.append(op.getValue()); gvar.append(op.getVariance())
                numObs += 1
                sigma2 = min(sigma2, op.getVariance())
            groups.append(dict(kind=kind, index=index, comp=comp, obs=obs, var=None if grp._packed is not None else gvar,
                               sigma=grp._packed, types=[op.getParameterType() for op in grp]))
        self._sigma2apriori = sigma2 if sigma2 > 0 else 1.0   # :221
        # ---- detectRankDefect -----------------------------------------------------------------------------------------
        hasBars = len(self._scaleBars) > 0
        free = dict(tx=True, ty=True, tz=True, rx=True, ry=True, rz=True, s=not hasBars)
        cnt = [0, 0, 0]
        X_T = (ParameterType.CAMERA_COORDINATE_X, ParameterType.OBJECT_COORDINATE_X)
        Y_T = (ParameterType.CAMERA_COORDINATE_Y, ParameterType.OBJECT_COORDINATE_Y)
        Z_T = (ParameterType.CAMERA_COORDINATE_Z, ParameterType.OBJECT_COORDINATE_Z)

        def none_free(): return not any(free.values())

        def rules():
            if free['tx'] and cnt[0] > 0: free['tx'] = False
            if free['ty'] and cnt[1] > 0: free['ty'] = False
            if free['tz'] and cnt[2] > 0: free['tz'] = False
            if not hasBars and (cnt[0] >= 2 or cnt[1] >= 2 or cnt[2] >= 2): free['s'] = False
            if free['rx'] and cnt[1] >= 2 and cnt[2] >= 2: free['rx'] = False
            if free['ry'] and cnt[0] >= 2 and cnt[2] >= 2: free['ry'] = False
            if free['rz'] and cnt[0] >= 2 and cnt[1] >= 2: free['rz'] = False
            if cnt[0] > 0 and cnt[1] > 0 and cnt[2] > 0 and sum(cnt) >= (6 if hasBars else 7):
                free['rx'] = free['ry'] = free['rz'] = False

        for g in groups:
            for t in g['types']:
                if t == ParameterType.CAMERA_OMEGA: free['rx'] = False
                elif t == ParameterType.CAMERA_PHI: free['ry'] = False
                elif t == ParameterType.CAMERA_KAPPA: free['rz'] = False
                if not free['rx'] and not free['ry'] and not free['rz']:
                    break
        for g in groups:
            for t in g['types']:
                if t in X_T: cnt[0] += 1
                elif t in Y_T: cnt[1] += 1
                elif t in Z_T: cnt[2] += 1
                elif t == ParameterType.CAMERA_OMEGA: free['rx'] = False
                elif t == ParameterType.CAMERA_PHI: free['ry'] = False
                elif t == ParameterType.CAMERA_KAPPA: free['rz'] = False
                rules()
                if none_free():
                    break
        done = False
        for g in oc_list:
            for c in range(3):
                cnt[c] += 1 if pt_col[g, c] == COL_FIXED else 0
            rules()
            if none_free():
                done = True
                break
        if not done and not none_free():
            for ep in eo_params:
                if free['rx'] and ep[3].getColumn() == COL_FIXED: free['rx'] = False
                if free['ry'] and ep[4].getColumn() == COL_FIXED: free['ry'] = False
                if free['rz'] and ep[5].getColumn() == COL_FIXED: free['rz'] = False
                for c in range(3):
                    cnt[c] += 1 if ep[c].getColumn() == COL_FIXED else 0
                rules()
                if none_free():
                    break
        free_flags = [int(free[k]) for k in ('tx', 'ty', 'tz', 'rx', 'ry', 'rz', 's')]
        d = sum(free_flags)
        # ---- renumber (:776-781) and write the columns back into the object graph ---------------------------------------
        for p in assigned:
            p.setColumn(seen[id(p)] + d)
        shift = (pt_col >= 0) & (pt_col != COL_FIXED)
        pt_col[shift] += d
        off = 0
        for s in stores:
            s.column[:] = pt_col[off:off + len(s)]
            off += len(s)
        self._numObs, self._numUnknown, self._defect = numObs, counter, d
        self._numIO, self._numDist, self._numObjectCoordinates = nIO, nDist, len(oc_list)
        self._objectCoordinates = [ObjectCoordinate._view(*self._locate(stores, g)) for g in oc_list] if len(oc_list) <= 100000 else []
        # ---- flatten ------------------------------------------------------------------------------------------------------
        io_val, io_col, r0, coef_ptr, ctype, cord, cval, ccol = [], [], [], [0], [], [], [], []
        for cam, plist in zip(self._cameras, cam_params):
            for p in plist[:3]:
                io_val.append(p.getValue()); io_col.append(p.getColumn())
            r0.append(cam._r0)
            for p in plist[3:]:
                ctype.append(int(p.getParameterType())); cord.append(p.getOrder() if isinstance(p, PolynomialCoefficient) else 0)
                cval.append(p.getValue()); ccol.append(p.getColumn())
            coef_ptr.append(len(ctype))
        eo_val = [p.getValue() for ep in eo_params for p in ep]
        eo_col = [p.getColumn() for ep in eo_params for p in ep]
        for g in groups:     # coefficient targets index the GLOBAL coefficient list in the C ABI
            for i, k in enumerate(g['kind']):
                if k == 2:
                    g['index'][i] = coef_ptr[g['index'][i]] + g['comp'][i]
                    g['comp'][i] = 0
        xyz = np.concatenate([s.xyz for s in stores]) if stores else np.zeros((0, 3))
        flat = dict(
            io_val=np.array(io_val, float), io_col=np.array(io_col, np.int64), r0=np.array(r0, float),
            coef_ptr=np.array(coef_ptr, np.int32), coef_type=np.array(ctype, np.int32), coef_order=np.array(cord, np.int32),
            coef_val=np.array(cval, float), coef_col=np.array(ccol, np.int64),
            cam_of_img=np.array(cam_of_img, np.int32), eo_val=np.array(eo_val, float), eo_col=np.array(eo_col, np.int64),
            pt_ptr=np.array(pt_ptr, np.int64), obj_idx=obj.astype(np.int32),
            xy=(np.concatenate(xy) if xy else np.zeros((0, 2))).reshape(-1),
            var=(np.concatenate(var) if var else np.zeros((0, 2))).reshape(-1),
            rho=np.concatenate(rho) if rho else np.zeros(0),
            xyz=xyz.reshape(-1).copy(), pt_col=pt_col.reshape(-1), is_datum=(datum & in_adjustment).astype(np.uint8),
            bar_a=np.array(bar_a, np.int32), bar_b=np.array(bar_b, np.int32), bar_len=np.array(bar_len, float),
            bar_var=np.array(bar_var, float), groups=groups, free_flags=np.array(free_flags, np.int32),
            n_unknowns=counter, n_observations=numObs, img_sigma=img_sigma)
        self._flat_ctx = (stores, cam_params, eo_params)
        self._store_base, self._image_index = dict(base), {id(img): k for k, img in enumerate(images)}
        return flat

    @staticmethod
    def _locate(stores, g):
        for s in stores:
            if g < len(s):
                return s, g
            g -= len(s)
        raise IndexError(g)

    def _fire(self, state, old, new):
        for l in self._listeners:
            l(state, old, new)

    # ---- estimateModel, :203-387 --------------------------------------------------------------------------------------
    def estimateModel(self):
        flat = self._prepare()
        # numRows of the reduced system, :262
        flat['reduced_rows'] = self._numIO + self._numDist + 3 * self._numObjectCoordinates + self._defect
        self._Qxx = None
        self._session = _lib.Session(
            invert_mode={MatrixInversion.NONE: _lib.INVERT_NONE, MatrixInversion.FULL: _lib.INVERT_FULL,
                         MatrixInversion.PRE_ELIMINATION: _lib.INVERT_PRE_ELIMINATION,
                         MatrixInversion.REDUCED: _lib.INVERT_REDUCED}[self._invert],
            estimation_type=_lib.SIMULATION if self._estimationType == EstimationType.SIMULATION else _lib.L2NORM,
            max_iterations=self._maxIter, use_centroid=self._useCentroid, apply_aposteriori=self._applyAposteriori,
            device=self._device, sigma2apriori=self._sigma2apriori, damping_value=self._damping,
            solver=getattr(self, '_solver', _lib.SOLVER_AUTO))
        self._session.set_problem(flat)
        if getattr(self, '_interrupt_requested', False):
            self._session.interrupt()
        self._interrupt_requested = False
        rc = self._session.estimate(progress=self._fire if self._listeners else None)
        self._interrupt_requested = False
        self.stats = self._session.stats()
        self._omega = self.stats.omega
        # write the adjusted values back into the object graph
        stores, cam_params, eo_params = self._flat_ctx
        xyz, io, coef, eo = self._session.values()
        off = 0
        xyz = xyz.reshape(-1, 3)
        for s in stores:
            s.xyz[:] = xyz[off:off + len(s)]
            off += len(s)
        ki = kc = 0
        for plist in cam_params:
            for p in plist[:3]:
                p.setValue(io[ki]); ki += 1
            for p in plist[3:]:
                p.setValue(coef[kc]); kc += 1
        ke = 0
        for ep in eo_params:
            for p in ep:
                p.setValue(eo[ke]); ke += 1
        self._status = EstimationStateType(rc) if rc in (1, -1, -2, -4, -7) else EstimationStateType.NOT_INITIALISED
        if rc in (1, -4) and self._writer is not None:                # exportAdjustmentResults, :360-368, :1164-1171
            try:
                self._writer.export(self)
            except (ValueError, TypeError, OSError):
                self._status = EstimationStateType.EXPORT_ADJUSTMENT_RESULTS_FAILED
        return self._status
