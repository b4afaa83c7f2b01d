"""jaicov_b200 -- B200-native (sm_100a) least-squares adjustment hot path for JAICOV
(applied-geodesy/bundle-adjustment).

* ``csrc/``      hand-written CUDA kernels + the C ABI (``include/jaicov_b200.h``) -> ``libjaicov_b200.so``
* ``_lib``       ctypes binding of the C ABI (``Session``)
* ``workloads``  synthetic networks of the benchmark configurations + scene -> object graph -> flat arrays
* ``verify``     size-independent residual checks of a final pass (matrix-free K x on the device)
* ``host``       host-side mirror of the JAICOV API (Camera, Image, BundleAdjustment, ...): object graph, integer
                 bookkeeping and flattening -- the part that stays in Java in a real integration

There is no CPU compute path: importing works everywhere, computing needs the built library and a B200.
"""
from . import _lib
from ._lib import JaicovError, Session, spd_solve_invert
from .host import (AffinityShearDistortionModel, BundleAdjustment, Camera, CoordinateTransformationExteriorOrientation,
                   DirectLinearTransformation, DirectlyObservedParameterGroup, DistortionModel, DLTCoefficients,
                   EstimationStateType, EstimationType, ExteriorOrientation, Image, InteriorOrientation, MatrixInversion,
                   ObjectCoordinate, ObjectCoordinateArray, ObservationParameter, ParameterType, PolynomialCoefficient,
                   RadialDistanceDistortionModel, RadiallySymmetricDistortionModel, ScaleBar, TangentialDistortionModel,
                   UnknownParameter, UpperSymmPackMatrix, ZernikeDistortionModel)

from .writers import DefaultResultWriter, MatlabResultWriter
from . import verify, workloads

__all__ = [n for n in dir() if not n.startswith('_')]
