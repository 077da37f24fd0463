"""topopt_jl_b200 -- host-side mirror of TopOpt.jl's SIMP inner-loop interface over
libtopopt_cuda (hand-written fp64 CUDA for sm_100a).  See DESIGN.md / INTEGRATION.md."""
from . import _lib
from .cheqfilters import DensityFilterFun, SensFilterFun
from .fea import (
    AbstractLinearSolver,
    CUDAAssemblySolver,
    CUDAMatrixFreeSolver,
    DefaultCriteria,
    EnergyCriteria,
    FEASolver,
    GenericFEASolver,
    HeavisideProjectionFun,
    PowerPenaltyFun,
    ProjectedPenaltyFun,
    PseudoDensities,
    RationalPenaltyFun,
    SigmoidProjectionFun,
    SinhPenaltyFun,
    getcompliance,
)
from .io import device_design, save_mesh
from .functions import ComplianceFun, DisplacementFun, TemperatureFun, ThermalComplianceFun, VolumeFun
from .problems import HalfMBB, HeatConductionProblem, HeatTree, Metadata, PointLoadCantilever, element_matrix
from .simp import oc_update, oc_update_device, simp_eval, simp_loop, simp_loop_device

__all__ = [n for n in dir() if not n.startswith("_")]
