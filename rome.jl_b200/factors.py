"""Host-side mirror of the reference's factor / variable / belief types for the hot path.

Same names and argument meaning as RoME.jl so graph-building code reads like the reference:
    Pose2Pose2(MvNormal([10, 0, pi/3], diagm([0.1, 0.1, 0.1].^2)))       src/factors/Pose2D.jl:30-32
    PriorPose2(MvNormal(zeros(3), 0.01*I))                               src/factors/PriorPose2.jl:13-15
    Pose2Point2BearingRange(Normal(0, 0.1), Normal(20, 1))               src/factors/BearingRange2D.jl:10-13
    Pose3Pose3(MvNormal(zeros(6), diagm([0.01*ones(3); 1e-4*ones(3)])))  src/factors/Pose3Pose3.jl:9-11
    PriorPose3(...)                                                      src/factors/Pose3D.jl:9-11
These objects only HOLD parameters; every evaluation goes to the CUDA library.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from . import _lib as L


# ---- beliefs (Distributions.jl names) ------------------------------------------------------------
@dataclass
class Normal:
    mu: float = 0.0
    sigma: float = 1.0

    def __post_init__(self):
        if not self.sigma > 0:
            raise ValueError("Normal: sigma must be positive")


@dataclass(init=False)
class MvNormal:
    mu: np.ndarray
    Sigma: np.ndarray

    def __init__(self, mu, Sigma=None):
        mu = np.atleast_1d(np.asarray(mu, dtype=np.float64))
        if Sigma is None:  # MvNormal(Diagonal) form: MvNormal(diag_cov_matrix)
            Sigma, mu = np.asarray(mu), np.zeros(np.asarray(mu).shape[0])
        Sigma = np.asarray(Sigma, dtype=np.float64)
        if Sigma.ndim == 1:  # vector of standard deviations, Distributions.MvNormal(mu, sigma_vec)
            Sigma = np.diag(Sigma ** 2)
        if Sigma.shape != (mu.shape[0], mu.shape[0]):
            raise ValueError("MvNormal: covariance shape does not match the mean")
        try:
            np.linalg.cholesky(Sigma)
        except np.linalg.LinAlgError as e:  # Distributions throws PosDefException
            raise ValueError("MvNormal: covariance is not positive definite") from e
        self.mu, self.Sigma = mu, Sigma


# ---- variable types (src/variables/VariableTypes.jl:13,35,47) --------------------------------------
class InferenceVariable:
    vartype: int
    dim: int
    manifold: str


class Pose2(InferenceVariable):
    vartype, dim, manifold = L.POSE2, 3, "SpecialEuclidean(2; vectors=HybridTangentRepresentation())"


class Point2(InferenceVariable):
    vartype, dim, manifold = L.POINT2, 2, "TranslationGroup(2)"


class Pose3(InferenceVariable):
    vartype, dim, manifold = L.POSE3, 6, "SpecialEuclidean(3; vectors=HybridTangentRepresentation())"


class Point3(InferenceVariable):  # src/variables/VariableTypes.jl:23
    vartype, dim, manifold = L.POINT3, 3, "TranslationGroup(3)"


class Rotation3(InferenceVariable):  # src/variables/VariableTypes.jl:50; coordinates = rotation vector
    vartype, dim, manifold = L.ROTATION3, 3, "SpecialOrthogonal(3)"


# ---- factors -----------------------------------------------------------------------------------------
class AbstractFactor:
    family: int
    variabletypes: tuple
    is_prior = False


class AbstractManifoldMinimize(AbstractFactor):
    pass


class AbstractPrior(AbstractFactor):
    is_prior = True


def _default_mv(d, diag):
    return MvNormal(np.zeros(d), np.diag(diag))


@dataclass
class Pose2Pose2(AbstractManifoldMinimize):
    Z: MvNormal = field(default_factory=lambda: _default_mv(3, [1.0, 1.0, 1.0]))
    family = L.POSE2POSE2
    variabletypes = (Pose2, Pose2)

    def __post_init__(self):
        if self.Z.mu.shape[0] != 3:
            raise ValueError("Pose2Pose2 needs a 3-dimensional belief")


@dataclass
class PriorPose2(AbstractPrior):
    Z: MvNormal = field(default_factory=lambda: _default_mv(3, [1.0, 1.0, 0.1]))
    family = L.PRIORPOSE2
    variabletypes = (Pose2,)

    def __post_init__(self):
        if self.Z.mu.shape[0] != 3:
            raise ValueError("PriorPose2 needs a 3-dimensional belief")


@dataclass
class Pose2Point2BearingRange(AbstractManifoldMinimize):
    bearing: Normal
    range: Normal
    family = L.BEARINGRANGE
    variabletypes = (Pose2, Point2)


@dataclass
class Pose3Pose3(AbstractManifoldMinimize):
    Z: MvNormal = field(default_factory=lambda: _default_mv(6, [0.01] * 3 + [0.0001] * 3))
    family = L.POSE3POSE3
    variabletypes = (Pose3, Pose3)

    def __post_init__(self):
        if self.Z.mu.shape[0] != 6:
            raise ValueError("Pose3Pose3 needs a 6-dimensional belief")


@dataclass
class PriorPose3(AbstractPrior):
    Z: MvNormal = field(default_factory=lambda: _default_mv(6, [0.01] * 3 + [0.0001] * 3))
    family = L.PRIORPOSE3
    variabletypes = (Pose3,)

    def __post_init__(self):
        if self.Z.mu.shape[0] != 6:
            raise ValueError("PriorPose3 needs a 6-dimensional belief")


# ---- next-row families --------------------------------------------------------------
@dataclass
class PriorPoint2(AbstractPrior):  # src/factors/Point2D.jl:8-10
    Z: MvNormal = field(default_factory=lambda: _default_mv(2, [0.01, 0.01]))
    family = L.PRIORPOINT2
    variabletypes = (Point2,)


@dataclass
class Point2Point2(AbstractManifoldMinimize):  # src/factors/Point2D.jl:25-27
    Z: MvNormal = field(default_factory=lambda: _default_mv(2, [0.1, 0.1]))
    family = L.POINT2POINT2
    variabletypes = (Point2, Point2)


@dataclass
class Pose2Point2(AbstractManifoldMinimize):  # src/factors/Pose2Point2.jl:9-12
    Z: MvNormal = field(default_factory=lambda: _default_mv(2, [0.01, 0.01]))
    family = L.POSE2POINT2
    variabletypes = (Pose2, Point2)


@dataclass
class Pose2Point2Range(AbstractManifoldMinimize):  # src/factors/Range2D.jl:43-47
    Z: Normal
    family = L.POSE2POINT2RANGE
    variabletypes = (Pose2, Point2)


@dataclass
class Point2Point2Range(AbstractManifoldMinimize):  # src/factors/Range2D.jl:5-7
    Z: Normal
    family = L.POINT2POINT2RANGE
    variabletypes = (Point2, Point2)


@dataclass
class Pose2Point2Bearing(AbstractManifoldMinimize):  # src/factors/Bearing2D.jl:13-15
    Z: Normal = field(default_factory=Normal)
    family = L.POSE2POINT2BEARING
    variabletypes = (Pose2, Point2)


@dataclass
class PriorPoint3(AbstractPrior):  # src/factors/Point3D.jl:7-9
    Z: MvNormal = field(default_factory=lambda: _default_mv(3, [1.0, 1.0, 1.0]))
    family = L.PRIORPOINT3
    variabletypes = (Point3,)


@dataclass
class Point3Point3(AbstractManifoldMinimize):  # src/factors/Point3Point3.jl:4-6
    Z: MvNormal = field(default_factory=lambda: _default_mv(3, [0.1, 0.1, 0.1]))
    family = L.POINT3POINT3
    variabletypes = (Point3, Point3)


@dataclass
class Pose3Pose3XYYaw(AbstractManifoldMinimize):  # src/factors/PartialPose3.jl:103-113, partial = (1, 2, 6)
    Z: MvNormal
    family = L.POSE3POSE3XYYAW
    variabletypes = (Pose3, Pose3)
    partial = (1, 2, 6)

    def __post_init__(self):
        if self.Z.mu.shape[0] != 3:
            raise ValueError("Pose3Pose3XYYaw needs a 3-dimensional (x, y, yaw) belief")


@dataclass
class Pose3Pose3Rotation(AbstractManifoldMinimize):  # src/factors/PartialPose3.jl:198-204, partial = (4, 5, 6)
    Z: MvNormal
    family = L.POSE3POSE3ROTATION
    variabletypes = (Pose3, Pose3)
    partial = (4, 5, 6)

    def __post_init__(self):
        if self.Z.mu.shape[0] != 3:
            raise ValueError("Pose3Pose3Rotation needs a 3-dimensional rotation-vector belief")


@dataclass
class Pose3Pose3UnitTrans(AbstractManifoldMinimize):  # src/factors/Pose3Pose3.jl:100-105
    Z: MvNormal = field(default_factory=lambda: _default_mv(6, [0.01] * 3 + [0.0001] * 3))
    family = L.POSE3POSE3UNITTRANS
    variabletypes = (Pose3, Pose3)


@dataclass
class Pose3Pose3RotOffset(AbstractManifoldMinimize):  # src/factors/Pose3Pose3.jl:57-78
    """odometry measured in a rotated frame: qhat = p o (m.t, bRa Exp(m.w)) with bRa a Rotation3 variable"""
    Z: MvNormal = field(default_factory=lambda: _default_mv(6, [0.01] * 3 + [0.0001] * 3))
    family = L.POSE3POSE3ROTOFFSET
    variabletypes = (Pose3, Pose3, Rotation3)


@dataclass
class Pose3Pose3Transform(AbstractManifoldMinimize):  # src/factors/Pose3Pose3.jl:80-95
    """qhat = p o Delta o exp(m) with the extrinsic Delta a Pose3 variable"""
    Z: MvNormal = field(default_factory=lambda: _default_mv(6, [0.01] * 3 + [0.0001] * 3))
    family = L.POSE3POSE3TRANSFORM
    variabletypes = (Pose3, Pose3, Pose3)


TERNARY_FACTORS = (Pose3Pose3RotOffset, Pose3Pose3Transform)
SCALAR_FACTORS = (Pose2Point2Range, Point2Point2Range, Pose2Point2Bearing)
POINT2_FACTORS = (PriorPoint2, Point2Point2, Pose2Point2)
PARTIAL_FACTORS = (Pose3Pose3XYYaw, Pose3Pose3Rotation, Pose3Pose3UnitTrans)  # no closed-form full proposal


def getManifold(x) -> str:
    """DFG.getManifold for the types of this path (returned as the reference's constructor text)."""
    if isinstance(x, type):
        x = x()
    if isinstance(x, (Pose2Pose2, PriorPose2, Pose2)):
        return Pose2.manifold  # src/factors/Pose2D.jl:34, PriorPose2.jl:17
    if isinstance(x, (Pose3Pose3, PriorPose3)) or x is Pose3 or type(x) is Pose3:
        return Pose3.manifold  # src/factors/Pose3Pose3.jl:13, Pose3D.jl:13
    if isinstance(x, Pose2Point2BearingRange):  # src/factors/BearingRange2D.jl:15
        return "ProductGroup(ProductManifold(SpecialOrthogonal(2), TranslationGroup(1)), LeftInvariantRepresentation())"
    if isinstance(x, (Point2, PriorPoint2, Point2Point2, Pose2Point2)):
        return Point2.manifold  # Point2D.jl:12,29; Pose2Point2.jl:18
    if isinstance(x, (Point3, PriorPoint3, Point3Point3)):
        return Point3.manifold  # Point3D.jl:11, Point3Point3.jl:8
    if isinstance(x, Pose3Pose3XYYaw):
        return Pose2.manifold  # PartialPose3.jl:113
    if isinstance(x, Pose3Pose3Rotation):
        return "SpecialOrthogonal(3)"  # PartialPose3.jl:204
    if isinstance(x, (Pose3Pose3UnitTrans, Pose3Pose3RotOffset, Pose3Pose3Transform)):
        return Pose3.manifold  # Pose3Pose3.jl:105, :61, :84
    if isinstance(x, Rotation3) or x is Rotation3:
        return Rotation3.manifold
    if isinstance(x, (Pose2Point2Range, Point2Point2Range)):
        return "TranslationGroup(1)"  # Range2D.jl:9,49
    if isinstance(x, Pose2Point2Bearing):
        return "SpecialOrthogonal(2)"  # Bearing2D.jl:19
    raise TypeError(f"no manifold for {type(x)}")


def getMeasurementParametric(f):
    """IIF.getMeasurementParametric: (mean, information).  BearingRange: src/factors/BearingRange2D.jl:30-37."""
    if isinstance(f, Pose2Point2BearingRange):
        return (np.array([f.bearing.mu, f.range.mu]),
                np.diag([1.0 / f.bearing.sigma ** 2, 1.0 / f.range.sigma ** 2]))
    if isinstance(f, SCALAR_FACTORS):
        return np.array([f.Z.mu]), np.array([[1.0 / f.Z.sigma ** 2]])
    return f.Z.mu.copy(), np.linalg.inv(f.Z.Sigma)


def factor_mean(f) -> np.ndarray:
    if isinstance(f, Pose2Point2BearingRange):
        return np.array([f.bearing.mu, f.range.mu])
    if isinstance(f, SCALAR_FACTORS):
        return np.array([f.Z.mu])
    return f.Z.mu


# ---- Packed* serialization types (Pose2D.jl:76-84, PriorPose2.jl:55-63, BearingRange2D.jl:76-87,
#      Pose3Pose3.jl:46-54, Pose3D.jl:28-36): belief fields as plain dictionaries ----------------------
def _pack_belief(b):
    if isinstance(b, Normal):
        return {"_type": "IncrementalInference.PackedNormal", "mu": b.mu, "sigma": b.sigma}
    return {"_type": "IncrementalInference.PackedFullNormal", "mu": b.mu.tolist(), "cov": b.Sigma.reshape(-1).tolist()}


def _unpack_belief(d):
    if d["_type"].endswith("PackedNormal"):
        return Normal(d["mu"], d["sigma"])
    mu = np.asarray(d["mu"], dtype=np.float64)
    return MvNormal(mu, np.asarray(d["cov"], dtype=np.float64).reshape(len(mu), len(mu)))


def pack(f) -> dict:
    """convert(PackedF, f)"""
    if isinstance(f, Pose2Point2BearingRange):
        return {"_type": "PackedPose2Point2BearingRange", "bearstr": _pack_belief(f.bearing),
                "rangstr": _pack_belief(f.range)}
    return {"_type": "Packed" + type(f).__name__, "Z": _pack_belief(f.Z)}


def unpack(d: dict):
    """convert(F, packed)"""
    name = d["_type"][len("Packed"):]
    if name == "Pose2Point2BearingRange":
        return Pose2Point2BearingRange(_unpack_belief(d["bearstr"]), _unpack_belief(d["rangstr"]))
    cls = {"Pose2Pose2": Pose2Pose2, "PriorPose2": PriorPose2, "Pose3Pose3": Pose3Pose3, "PriorPose3": PriorPose3,
           "PriorPoint2": PriorPoint2, "Point2Point2": Point2Point2, "Pose2Point2": Pose2Point2,
           "Pose2Point2Range": Pose2Point2Range, "Point2Point2Range": Point2Point2Range,
           "Pose2Point2Bearing": Pose2Point2Bearing, "PriorPoint3": PriorPoint3, "Point3Point3": Point3Point3,
           "Pose3Pose3XYYaw": Pose3Pose3XYYaw, "Pose3Pose3Rotation": Pose3Pose3Rotation,
           "Pose3Pose3UnitTrans": Pose3Pose3UnitTrans}[name]
    return cls(_unpack_belief(d["Z"]))
