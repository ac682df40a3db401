"""ctypes mirror of include/folp_b200.h (the C ABI of libfolp_b200.so).

Field order and types must match the header exactly; tests/test_abi.py checks
sizeof() of every struct against the values compiled into the library.
"""
import ctypes as C
import enum


class Status(enum.IntEnum):
    OK = 0
    INVALID_ARGUMENT = 1
    CUDA_ERROR = 2
    NCCL_ERROR = 3
    OUT_OF_MEMORY = 4
    UNSUPPORTED = 5


# --- enums with the Julia @enum ordinals (solve_log.jl, saddle_point.jl, termination.jl)
class RestartChoice(enum.IntEnum):
    RESTART_CHOICE_UNSPECIFIED = 0
    RESTART_CHOICE_NO_RESTART = 1
    RESTART_CHOICE_WEIGHTED_AVERAGE_RESET = 2
    RESTART_CHOICE_RESTART_TO_AVERAGE = 3


class PointType(enum.IntEnum):
    POINT_TYPE_UNSPECIFIED = 0
    POINT_TYPE_CURRENT_ITERATE = 1
    POINT_TYPE_ITERATE_DIFFERENCE = 2
    POINT_TYPE_AVERAGE_ITERATE = 3
    POINT_TYPE_NONE = 4


class TerminationReason(enum.IntEnum):
    TERMINATION_REASON_UNSPECIFIED = 0
    TERMINATION_REASON_OPTIMAL = 1
    TERMINATION_REASON_PRIMAL_INFEASIBLE = 2
    TERMINATION_REASON_DUAL_INFEASIBLE = 3
    TERMINATION_REASON_TIME_LIMIT = 4
    TERMINATION_REASON_ITERATION_LIMIT = 5
    TERMINATION_REASON_KKT_MATRIX_PASS_LIMIT = 6
    TERMINATION_REASON_NUMERICAL_ERROR = 7
    TERMINATION_REASON_INVALID_PROBLEM = 8
    TERMINATION_REASON_OTHER = 9


class RestartScheme(enum.IntEnum):
    NO_RESTARTS = 0
    FIXED_FREQUENCY = 1
    ADAPTIVE_NORMALIZED = 2
    ADAPTIVE_LOCALIZED = 3
    ADAPTIVE_DISTANCE = 4


class RestartToCurrentMetric(enum.IntEnum):
    NO_RESTART_TO_CURRENT = 0
    GAP_OVER_DISTANCE = 1
    GAP_OVER_DISTANCE_SQUARED = 2


class OptimalityNorm(enum.IntEnum):
    L_INF = 0
    L2 = 1


class StepSizePolicy(enum.IntEnum):
    ADAPTIVE = 0
    MALITSKY_POCK = 1
    CONSTANT = 2


_pd = C.POINTER(C.c_double)
_pi = C.POINTER(C.c_int64)


class FolpProblem(C.Structure):
    _fields_ = [
        ("num_variables", C.c_int64),
        ("num_constraints", C.c_int64),
        ("num_nonzeros", C.c_int64),
        ("num_equalities", C.c_int64),
        ("index_base", C.c_int32),
        ("reserved0", C.c_int32),
        ("colptr", _pi),
        ("rowval", _pi),
        ("nzval", _pd),
        ("objective_vector", _pd),
        ("variable_lower_bound", _pd),
        ("variable_upper_bound", _pd),
        ("right_hand_side", _pd),
        ("objective_constant", C.c_double),
        ("variable_rescaling", _pd),
        ("constraint_rescaling", _pd),
        ("orig_objective_vector", _pd),
        ("orig_variable_lower_bound", _pd),
        ("orig_variable_upper_bound", _pd),
        ("orig_right_hand_side", _pd),
        ("orig_nzval", _pd),
        ("q_num_nonzeros", C.c_int64),
        ("q_colptr", _pi),
        ("q_rowval", _pi),
        ("q_nzval", _pd),
        ("q_orig_nzval", _pd),
        ("l_inf_norm_primal_linear_objective", C.c_double),
        ("l_inf_norm_primal_right_hand_side", C.c_double),
        ("l2_norm_primal_linear_objective", C.c_double),
        ("l2_norm_primal_right_hand_side", C.c_double),
    ]


class FolpParams(C.Structure):
    _fields_ = [
        ("step_size_policy", C.c_int32),
        ("termination_evaluation_frequency", C.c_int32),
        ("reduction_exponent", C.c_double),
        ("growth_exponent", C.c_double),
        ("downscaling_factor", C.c_double),
        ("breaking_factor", C.c_double),
        ("interpolation_coefficient", C.c_double),
        ("initial_step_size", C.c_double),
        ("initial_primal_weight", C.c_double),
        ("initial_kkt_passes", C.c_double),
        ("optimality_norm", C.c_int32),
        ("iteration_limit", C.c_int32),
        ("eps_optimal_absolute", C.c_double),
        ("eps_optimal_relative", C.c_double),
        ("eps_primal_infeasible", C.c_double),
        ("eps_dual_infeasible", C.c_double),
        ("time_sec_limit", C.c_double),
        ("kkt_matrix_pass_limit", C.c_double),
        ("restart_scheme", C.c_int32),
        ("restart_to_current_metric", C.c_int32),
        ("restart_frequency_if_fixed", C.c_int64),
        ("artificial_restart_threshold", C.c_double),
        ("sufficient_reduction_for_restart", C.c_double),
        ("necessary_reduction_for_restart", C.c_double),
        ("primal_weight_update_smoothing", C.c_double),
        ("use_approximate_localized_duality_gap", C.c_int32),
        ("record_iteration_stats", C.c_int32),
        ("verbosity", C.c_int32),
        ("reserved0", C.c_int32),
    ]


CONVERGENCE_FIELDS = (
    "primal_objective",
    "dual_objective",
    "corrected_dual_objective",
    "l_inf_primal_residual",
    "l2_primal_residual",
    "l_inf_dual_residual",
    "l2_dual_residual",
    "relative_l_inf_primal_residual",
    "relative_l2_primal_residual",
    "relative_l_inf_dual_residual",
    "relative_l2_dual_residual",
    "relative_optimality_gap",
    "l_inf_primal_variable",
    "l2_primal_variable",
    "l_inf_dual_variable",
    "l2_dual_variable",
)
INFEASIBILITY_FIELDS = (
    "max_primal_ray_infeasibility",
    "primal_ray_linear_objective",
    "primal_ray_quadratic_norm",
    "max_dual_ray_infeasibility",
    "dual_ray_objective",
)


class FolpEval(C.Structure):
    _fields_ = (
        [("iteration_number", C.c_int32), ("candidate_type", C.c_int32)]
        + [(f, C.c_double) for f in CONVERGENCE_FIELDS]
        + [(f, C.c_double) for f in INFEASIBILITY_FIELDS]
        + [
            ("cumulative_kkt_matrix_passes", C.c_double),
            ("cumulative_time_sec", C.c_double),
            ("step_size", C.c_double),
            ("primal_weight", C.c_double),
            ("time_spent_doing_basic_algorithm", C.c_double),
            ("lagrangian_value", C.c_double),
            ("estimated_lower_bound", C.c_double),
            ("estimated_upper_bound", C.c_double),
            ("cumulative_rejected_steps", C.c_int32),
            ("restart_used", C.c_int32),
            ("termination_reason", C.c_int32),
            ("numerical_error", C.c_int32),
            ("total_number_iterations", C.c_int64),
        ]
    )

    def as_dict(self):
        return {name: getattr(self, name) for name, _ in self._fields_}


class FolpDist(C.Structure):
    _fields_ = [
        ("rank", C.c_int32),
        ("world_size", C.c_int32),
        ("device", C.c_int32),
        ("reserved0", C.c_int32),
        ("nccl_unique_id", C.c_void_p),
    ]


class FolpDebugScalars(C.Structure):
    _fields_ = [
        ("step_size", C.c_double),
        ("primal_weight", C.c_double),
        ("cumulative_kkt_passes", C.c_double),
        ("sum_primal_solution_weights", C.c_double),
        ("sum_dual_solution_weights", C.c_double),
        ("total_number_iterations", C.c_int64),
        ("iterations_completed", C.c_int64),
        ("sum_primal_solutions_count", C.c_int64),
        ("sum_dual_solutions_count", C.c_int64),
        ("numerical_error", C.c_int32),
        ("reserved0", C.c_int32),
        ("last_interaction", C.c_double),
        ("last_movement", C.c_double),
    ]

    def as_dict(self):
        return {name: getattr(self, name) for name, _ in self._fields_}


# Every symbol include/folp_b200.h declares (tests check the .so exports all).
EXPORTED_SYMBOLS = (
    "folp_nccl_unique_id",
    "folp_create",
    "folp_run",
    "folp_solve",
    "folp_get_solution",
    "folp_debug_attempts",
    "folp_debug_state",
    "folp_debug_set_state",
    "folp_debug_spmv",
    "folp_counters",
    "folp_destroy",
    "folp_last_error",
    "folp_build_info",
)
