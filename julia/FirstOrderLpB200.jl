# FirstOrderLpB200.jl -- reference-side binding of libfolp_b200.so.
#
# UNVERIFIED: no Julia toolchain exists in the build image, so this file has been
# written against include/folp_b200.h but never executed. It is the stub a
# FirstOrderLp.jl maintainer would `include` after src/primal_dual_hybrid_gradient.jl
# to route `optimize(::PdhgParameters, ::QuadraticProgrammingProblem)` through the
# B200 library. Everything up to the while-loop (validate, cached norms,
# rescale_problem, initial step size, initial primal weight;
# src/primal_dual_hybrid_gradient.jl:786-859) stays in Julia, exactly as today.
#
# Usage:
#   include("julia/FirstOrderLpB200.jl")
#   out = FirstOrderLpB200.optimize_b200(params, qp)     # ::FirstOrderLp.SaddlePointOutput

module FirstOrderLpB200

import FirstOrderLp
using LinearAlgebra
using SparseArrays

const LIB = get(ENV, "FOLP_B200_LIB", "libfolp_b200.so")

# struct folp_problem (include/folp_b200.h:102-143)
struct FolpProblem
  num_variables::Int64
  num_constraints::Int64
  num_nonzeros::Int64
  num_equalities::Int64
  index_base::Int32
  reserved0::Int32
  colptr::Ptr{Int64}
  rowval::Ptr{Int64}
  nzval::Ptr{Float64}
  objective_vector::Ptr{Float64}
  variable_lower_bound::Ptr{Float64}
  variable_upper_bound::Ptr{Float64}
  right_hand_side::Ptr{Float64}
  objective_constant::Float64
  variable_rescaling::Ptr{Float64}
  constraint_rescaling::Ptr{Float64}
  orig_objective_vector::Ptr{Float64}
  orig_variable_lower_bound::Ptr{Float64}
  orig_variable_upper_bound::Ptr{Float64}
  orig_right_hand_side::Ptr{Float64}
  orig_nzval::Ptr{Float64}
  q_num_nonzeros::Int64
  q_colptr::Ptr{Int64}
  q_rowval::Ptr{Int64}
  q_nzval::Ptr{Float64}
  q_orig_nzval::Ptr{Float64}
  l_inf_norm_primal_linear_objective::Float64
  l_inf_norm_primal_right_hand_side::Float64
  l2_norm_primal_linear_objective::Float64
  l2_norm_primal_right_hand_side::Float64
end

# struct folp_params (include/folp_b200.h:151-183)
struct FolpParams
  step_size_policy::Int32
  termination_evaluation_frequency::Int32
  reduction_exponent::Float64
  growth_exponent::Float64
  downscaling_factor::Float64
  breaking_factor::Float64
  interpolation_coefficient::Float64
  initial_step_size::Float64
  initial_primal_weight::Float64
  initial_kkt_passes::Float64
  optimality_norm::Int32
  iteration_limit::Int32
  eps_optimal_absolute::Float64
  eps_optimal_relative::Float64
  eps_primal_infeasible::Float64
  eps_dual_infeasible::Float64
  time_sec_limit::Float64
  kkt_matrix_pass_limit::Float64
  restart_scheme::Int32
  restart_to_current_metric::Int32
  restart_frequency_if_fixed::Int64
  artificial_restart_threshold::Float64
  sufficient_reduction_for_restart::Float64
  necessary_reduction_for_restart::Float64
  primal_weight_update_smoothing::Float64
  use_approximate_localized_duality_gap::Int32
  record_iteration_stats::Int32
  verbosity::Int32
  reserved0::Int32
end

# struct folp_eval (include/folp_b200.h:188-229)
struct FolpEval
  iteration_number::Int32
  candidate_type::Int32
  primal_objective::Float64
  dual_objective::Float64
  corrected_dual_objective::Float64
  l_inf_primal_residual::Float64
  l2_primal_residual::Float64
  l_inf_dual_residual::Float64
  l2_dual_residual::Float64
  relative_l_inf_primal_residual::Float64
  relative_l2_primal_residual::Float64
  relative_l_inf_dual_residual::Float64
  relative_l2_dual_residual::Float64
  relative_optimality_gap::Float64
  l_inf_primal_variable::Float64
  l2_primal_variable::Float64
  l_inf_dual_variable::Float64
  l2_dual_variable::Float64
  max_primal_ray_infeasibility::Float64
  primal_ray_linear_objective::Float64
  primal_ray_quadratic_norm::Float64
  max_dual_ray_infeasibility::Float64
  dual_ray_objective::Float64
  cumulative_kkt_matrix_passes::Float64
  cumulative_time_sec::Float64
  step_size::Float64
  primal_weight::Float64
  time_spent_doing_basic_algorithm::Float64
  lagrangian_value::Float64
  estimated_lower_bound::Float64
  estimated_upper_bound::Float64
  cumulative_rejected_steps::Int32
  restart_used::Int32
  termination_reason::Int32
  numerical_error::Int32
  total_number_iterations::Int64
end

function check(rc::Cint, handle::Ptr{Cvoid})
  if rc != 0
    msg = unsafe_string(ccall((:folp_last_error, LIB), Cstring, (Ptr{Cvoid},), handle))
    error("libfolp_b200 (status $rc): $msg")   # the reference raises error() too
  end
end

function to_iteration_stats(e::FolpEval)
  ci = FirstOrderLp.ConvergenceInformation()
  ci.candidate_type = FirstOrderLp.PointType(e.candidate_type)
  for f in (:primal_objective, :dual_objective, :corrected_dual_objective,
            :l_inf_primal_residual, :l2_primal_residual, :l_inf_dual_residual,
            :l2_dual_residual, :relative_l_inf_primal_residual,
            :relative_l2_primal_residual, :relative_l_inf_dual_residual,
            :relative_l2_dual_residual, :relative_optimality_gap,
            :l_inf_primal_variable, :l2_primal_variable, :l_inf_dual_variable,
            :l2_dual_variable)
    setfield!(ci, f, getfield(e, f))
  end
  ii = FirstOrderLp.InfeasibilityInformation()
  ii.candidate_type = FirstOrderLp.PointType(e.candidate_type)
  for f in (:max_primal_ray_infeasibility, :primal_ray_linear_objective,
            :primal_ray_quadratic_norm, :max_dual_ray_infeasibility, :dual_ray_objective)
    setfield!(ii, f, getfield(e, f))
  end
  s = FirstOrderLp.IterationStats()
  s.iteration_number = e.iteration_number
  s.convergence_information = [ci]
  s.infeasibility_information = [ii]
  s.cumulative_kkt_matrix_passes = e.cumulative_kkt_matrix_passes
  s.cumulative_rejected_steps = e.cumulative_rejected_steps
  s.cumulative_time_sec = e.cumulative_time_sec
  s.restart_used = FirstOrderLp.RestartChoice(e.restart_used)
  s.step_size = e.step_size
  s.primal_weight = e.primal_weight
  s.method_specific_stats = Dict{AbstractString,Float64}(
    "time_spent_doing_basic_algorithm" => e.time_spent_doing_basic_algorithm,
    "lagrangian_value" => e.lagrangian_value,
    "estimated_lower_bound" => e.estimated_lower_bound,
    "estimated_upper_bound" => e.estimated_upper_bound,
  )
  return s
end

"""
Drop-in for `FirstOrderLp.rescale_problem` (src/preprocess.jl:631-687) computed on the
device by `folp_rescale_problem`: the copied problem's arrays are rescaled in place, the
cumulative rescaling vectors come back. Same arithmetic and summation order as the
reference (bit-identical for rows / columns of up to 2048 entries and alpha = 1).
"""
function rescale_problem_b200(
  l_inf_ruiz_iterations::Int64,
  l2_norm_rescaling::Bool,
  pock_chambolle_alpha::Union{Float64,Nothing},
  original_problem::FirstOrderLp.QuadraticProgrammingProblem,
)
  p = deepcopy(original_problem)
  A, Q = p.constraint_matrix, p.objective_matrix
  m, n = size(A)
  con = Vector{Float64}(undef, m)
  var = Vector{Float64}(undef, n)
  GC.@preserve p con var begin
    rc = ccall(
      (:folp_rescale_problem, LIB),
      Cint,
      (
        Int64, Int64, Int32, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Ptr{Int64}, Ptr{Int64}, Ptr{Float64},
        Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int32, Int32, Int32, Float64,
        Ptr{Float64}, Ptr{Float64},
      ),
      m, n, Int32(1), A.colptr, A.rowval, A.nzval, Q.colptr, Q.rowval, Q.nzval,
      p.objective_vector, p.variable_lower_bound, p.variable_upper_bound, p.right_hand_side,
      Int32(l_inf_ruiz_iterations), Int32(0), Int32(l2_norm_rescaling),
      pock_chambolle_alpha === nothing ? -1.0 : pock_chambolle_alpha, con, var,
    )
  end
  rc == 0 || error(unsafe_string(ccall((:folp_last_error, LIB), Cstring, (Ptr{Cvoid},), C_NULL)))
  return FirstOrderLp.ScaledQpProblem(original_problem, p, con, var)
end

"""
Drop-in for `FirstOrderLp.optimize(params::PdhgParameters, qp)`
(src/primal_dual_hybrid_gradient.jl:782): same arguments, same
`SaddlePointOutput`.
"""
function optimize_b200(
  params::FirstOrderLp.PdhgParameters,
  original_problem::FirstOrderLp.QuadraticProgrammingProblem;
  device_rescaling::Bool = false,
  device_ids::Vector{Int32} = Int32[],   # more than one entry: folp_create_multi drives all of them from this call
)
  # ---- host half, unchanged (pdhg.jl:786-859) ----
  FirstOrderLp.validate(original_problem)
  qp_cache = FirstOrderLp.cached_quadratic_program_info(original_problem)
  scaled_problem =
    device_rescaling ?
    rescale_problem_b200(
      params.l_inf_ruiz_iterations,
      params.l2_norm_rescaling,
      params.pock_chambolle_alpha,
      original_problem,
    ) :
    FirstOrderLp.rescale_problem(
      params.l_inf_ruiz_iterations,
      params.l2_norm_rescaling,
      params.pock_chambolle_alpha,
      params.verbosity,
      original_problem,
    )
  problem = scaled_problem.scaled_qp
  if params.primal_importance <= 0 || !isfinite(params.primal_importance)
    error("primal_importance must be positive and finite")
  end
  sp = params.step_size_policy_params
  policy, kkt0 = Int32(0), 0.5
  red, grow, down, brk, interp = 0.0, 0.0, 0.0, 0.0, 0.0
  if sp isa FirstOrderLp.AdaptiveStepsizeParams
    policy = Int32(0)
    red, grow = sp.reduction_exponent, sp.growth_exponent
    step0 = 1.0 / norm(problem.constraint_matrix, Inf)
  elseif sp isa FirstOrderLp.MalitskyPockStepsizeParameters
    policy = Int32(1)
    down, brk, interp = sp.downscaling_factor, sp.breaking_factor, sp.interpolation_coefficient
    step0 = 1.0 / norm(problem.constraint_matrix, Inf)
  else
    policy = Int32(2)
    sigma, iters = FirstOrderLp.estimate_maximum_singular_value(
      problem.constraint_matrix,
      probability_of_failure = 0.001,
      desired_relative_error = 0.2,
    )
    step0 = (1 - 0.2) / sigma
    kkt0 = Float64(iters)
  end
  n, m = length(problem.variable_lower_bound), length(problem.right_hand_side)
  pw0 = params.scale_invariant_initial_primal_weight ?
    FirstOrderLp.select_initial_primal_weight(problem, ones(n), ones(m),
                                              params.primal_importance, params.verbosity) :
    params.primal_importance

  tc, rp = params.termination_criteria, params.restart_params
  fparams = FolpParams(
    policy, Int32(params.termination_evaluation_frequency), red, grow, down, brk, interp,
    step0, pw0, kkt0,
    Int32(tc.optimality_norm), Int32(min(tc.iteration_limit, typemax(Int32))),
    tc.eps_optimal_absolute, tc.eps_optimal_relative, tc.eps_primal_infeasible,
    tc.eps_dual_infeasible, tc.time_sec_limit, tc.kkt_matrix_pass_limit,
    Int32(rp.restart_scheme), Int32(rp.restart_to_current_metric),
    Int64(rp.restart_frequency_if_fixed), rp.artificial_restart_threshold,
    rp.sufficient_reduction_for_restart, rp.necessary_reduction_for_restart,
    rp.primal_weight_update_smoothing, Int32(rp.use_approximate_localized_duality_gap),
    Int32(params.record_iteration_stats), Int32(params.verbosity), Int32(0),
  )

  A, Q = problem.constraint_matrix, problem.objective_matrix
  O = scaled_problem.original_qp
  n_out, m_out = n, m
  x_out, y_out = Vector{Float64}(undef, n_out), Vector{Float64}(undef, m_out)
  # every evaluation when record_iteration_stats is set (:958-960), the final record only otherwise
  max_evals = params.record_iteration_stats ?
    Int(min(tc.iteration_limit, typemax(Int32))) ÷ max(1, params.termination_evaluation_frequency) + 16 : 1
  evals = Vector{FolpEval}(undef, max_evals)
  num_evals, reason, iters = Ref{Int64}(0), Ref{Int32}(0), Ref{Int32}(0)
  handle = Ref{Ptr{Cvoid}}(C_NULL)
  # All arrays are borrowed for the duration of folp_create only.
  GC.@preserve A Q O problem scaled_problem begin
    fp = Ref(FolpProblem(
      n, m, nnz(A), problem.num_equalities, Int32(1), Int32(0),   # index_base = 1: Julia CSC as is
      pointer(A.colptr), pointer(A.rowval), pointer(A.nzval),
      pointer(problem.objective_vector), pointer(problem.variable_lower_bound),
      pointer(problem.variable_upper_bound), pointer(problem.right_hand_side),
      problem.objective_constant,
      pointer(scaled_problem.variable_rescaling), pointer(scaled_problem.constraint_rescaling),
      pointer(O.objective_vector), pointer(O.variable_lower_bound),
      pointer(O.variable_upper_bound), pointer(O.right_hand_side),
      Ptr{Float64}(C_NULL),                                       # orig_nzval: oracle only
      nnz(Q), pointer(Q.colptr), pointer(Q.rowval), pointer(Q.nzval),
      nnz(O.objective_matrix) == nnz(Q) ? pointer(O.objective_matrix.nzval) : Ptr{Float64}(C_NULL),
      qp_cache.l_inf_norm_primal_linear_objective, qp_cache.l_inf_norm_primal_right_hand_side,
      qp_cache.l2_norm_primal_linear_objective, qp_cache.l2_norm_primal_right_hand_side,
    ))
    if length(device_ids) > 1
      # one call, several GPUs of this process: threads of the library drive the devices (folp_create_multi)
      rc = ccall((:folp_create_multi, LIB), Cint,
                 (Ref{FolpProblem}, Ref{FolpParams}, Int32, Ptr{Int32}, Ref{Ptr{Cvoid}}),
                 fp, Ref(fparams), Int32(length(device_ids)), device_ids, handle)
    else
      rc = ccall((:folp_create, LIB), Cint,
                 (Ref{FolpProblem}, Ref{FolpParams}, Ptr{Cvoid}, Ref{Ptr{Cvoid}}),
                 fp, Ref(fparams), C_NULL, handle)
    end
    check(rc, C_NULL)
  end
  stats = FirstOrderLp.IterationStats[]
  try
    if params.verbosity <= 0
      # nothing to print: the whole loop in one C call
      rc = ccall((:folp_solve, LIB), Cint,
                 (Ptr{Cvoid}, Ptr{FolpEval}, Int64, Ref{Int64}, Ref{Int32}, Ref{Int32},
                  Ptr{Float64}, Ptr{Float64}),
                 handle[], evals, length(evals), num_evals, reason, iters, x_out, y_out)
      check(rc, handle[])
      num_evals[] > length(evals) && @warn "iteration statistics truncated" num_evals[] length(evals)
      stats = [to_iteration_stats(evals[k]) for k in 1:min(num_evals[], length(evals))]
    else
      # verbosity > 0: the loop of pdhg.jl:886-1048 evaluation by evaluation (folp_run), printing with the
      # reference's own functions where the reference prints (iteration table :962-970, final log :972-993)
      FirstOrderLp.display_iteration_stats_heading(params.verbosity)
      e = Ref{FolpEval}()
      while true
        rc = ccall((:folp_run, LIB), Cint, (Ptr{Cvoid}, Ref{FolpEval}), handle[], e)
        check(rc, handle[])
        st = to_iteration_stats(e[])
        iteration = Int(e[].iteration_number) + 1
        terminated = e[].termination_reason != 0
        (params.record_iteration_stats || terminated) && push!(stats, st)
        if FirstOrderLp.print_to_screen_this_iteration(
          terminated ? FirstOrderLp.TerminationReason(e[].termination_reason) : false,
          iteration, params.verbosity, params.termination_evaluation_frequency)
          FirstOrderLp.display_iteration_stats(st, params.verbosity)
        end
        if terminated
          reason[] = e[].termination_reason
          iters[] = e[].iteration_number
          rc = ccall((:folp_get_solution, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Float64}, Ptr{Float64}),
                     handle[], 0, 1, x_out, y_out)
          check(rc, handle[])
          xs, ys = similar(x_out), similar(y_out)   # the reference logs the SCALED average on the scaled problem
          rc = ccall((:folp_get_solution, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Float64}, Ptr{Float64}),
                     handle[], 0, 0, xs, ys)
          check(rc, handle[])
          FirstOrderLp.pdhg_final_log(problem, xs, ys, params.verbosity, iteration,
                                      FirstOrderLp.TerminationReason(reason[]), st)
          break
        end
      end
    end
  finally
    ccall((:folp_destroy, LIB), Cvoid, (Ptr{Cvoid},), handle[])
  end
  termination_reason = FirstOrderLp.TerminationReason(reason[])
  return FirstOrderLp.SaddlePointOutput(
    x_out, y_out, termination_reason,
    FirstOrderLp.termination_reason_to_string(termination_reason),
    iters[], stats,
  )
end

end  # module
