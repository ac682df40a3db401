"""folp_b200 -- B200-native PDHG inner loop behind FirstOrderLp.optimize.

Host-side mirror of the reference's Julia API for the PDHG path (problem types,
parameters, preprocessing, optimize(), solve-log records); the iteration loop
itself lives in csrc/ (hand-written sm_100a CUDA behind the C ABI declared in
include/folp_b200.h) and is reached through lib.py. There is no CPU fallback:
optimize() raises if libfolp_b200.so cannot be loaded.
"""
from ._abi import (  # noqa: F401
    OptimalityNorm,
    PointType,
    RestartChoice,
    RestartScheme,
    RestartToCurrentMetric,
    Status,
    StepSizePolicy,
    TerminationReason,
)
from .params import (  # noqa: F401
    AdaptiveStepsizeParams,
    ConstantStepsizeParams,
    MalitskyPockStepsizeParameters,
    PdhgParameters,
    RestartParameters,
    TerminationCriteria,
    construct_restart_parameters,
    construct_termination_criteria,
)
from .problem import (  # noqa: F401
    QuadraticProgrammingProblem,
    ScaledQpProblem,
    cached_quadratic_program_info,
    is_linear_programming_problem,
    linear_programming_problem,
    validate,
)
from .preprocess import (  # noqa: F401
    presolve,
    rescale_problem,
    undo_presolve,
)
from .solve_log import (  # noqa: F401
    ConvergenceInformation,
    InfeasibilityInformation,
    IterationStats,
    SaddlePointOutput,
    SolveLog,
    termination_reason_to_string,
)

from . import distributed  # noqa: F401
from .pdhg import estimate_maximum_singular_value, host_setup, optimize  # noqa: F401

__all__ = [n for n in dir() if not n.startswith("_")]
