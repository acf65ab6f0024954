"""rome_b200 -- B200 (sm_100a) implementation of RoME.jl's factor-residual / convolution hot path.

The directory is named `rome.jl_b200`; import it as `rome_b200` (see /rome_b200.py).
Compute lives in librome_b200.so (csrc/, hand-written CUDA behind the C ABI of
include/rome_b200.h); this package is the host-side mirror of the reference's factor API.
"""
from ._lib import (BARRIER_SIGNAL, BARRIER_WAIT, BEARINGRANGE, DECONV, INDEPENDENT, JACOBIAN, POINT2, POINT2POINT2, POINT2POINT2RANGE, POINT3, POINT3POINT3,
                   POSE2, POSE2POINT2, POSE2POINT2BEARING, POSE2POINT2RANGE, POSE2POSE2, POSE3, POSE3POSE3,
                   POSE3POSE3ROTATION, POSE3POSE3ROTOFFSET, POSE3POSE3TRANSFORM, POSE3POSE3UNITTRANS, POSE3POSE3XYYAW, PRECISE, PRIORPOINT2, PRIORPOINT3, PRIORPOSE2, PRIORPOSE3,
                   PROPOSAL_BWD, PROPOSAL_FWD, RESIDUAL, ROTATION3, ROUTED_ONLY, SAMPLE, SO_PATH, STATS, SYMBOLS, WRITE_MEAS, RomeB200Error)
from .engine import (BYTES_PER_EVAL, BYTES_PER_EVAL_SAMPLED, FAMILY, FAMILY_VT2, VAR_DIM, Context, dequantized_particles,
                     meas_to_offsets, npad, offsets_to_meas, plan_query, rows_to_particle_major)

from .factors import (MvNormal, Normal, Point2, Point2Point2, Point2Point2Range, Point3, Point3Point3, Pose2, Pose2Point2,
                      Pose2Point2Bearing, Pose2Point2BearingRange, Pose2Point2Range, Pose2Pose2, Pose3, Pose3Pose3,
                      Pose3Pose3Rotation, Pose3Pose3RotOffset, Pose3Pose3Transform, Pose3Pose3UnitTrans, Pose3Pose3XYYaw, PriorPoint2,
                      PriorPoint3, PriorPose2, Rotation3,
                      PriorPose3, getManifold, getMeasurementParametric, pack, unpack)
from .graph import (DeviceGraph, FactorGraph, SolverParams, accumulateFactorMeans, addFactor, addVariable, approxConv, approxConvBelief, approxDeconv,
                    calcFactorResidual, calcFactorResidualTemporary, default_context, getSample, getSolverParams,
                    getVal, initAll, initfg, ls, lsf, sampleFactor, setVal, solveFactorParametric)
from .canonical import (generateGraph_Beehive, generateGraph_Circle, generateGraph_Hexagonal, generateGraph_Honeycomb,
                        generateGraph_ManhattanShaped, generateGraph_Pose3Chain, generateGraph_ZeroPose,
                        seed_particles)
from . import sharding
from .solver import GibbsSolver, build_product_plans, solveGraphGibbs
from .parametric import color_variables, solveGraphParametric
from .dfg_io import loadDFG, saveDFG, setPPE
from .g2o import exportG2o, graphFromEdgeArrays, importG2o, loadG2o, parseG2oInstruction, stringG2o

__version__ = "0.1.0"
