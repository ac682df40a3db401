/*
 * folp_oracle.c -- CPU oracle: plain-C restatement of the PDHG path of
 * google-research/FirstOrderLp.jl.  TEST INFRASTRUCTURE ONLY (see
 * folp_oracle.h).  Build with -ffp-contract=off so that no multiply-add is
 * fused: the Julia reference does not contract.
 *
 * Abbreviations in citations:
 *   pdhg.jl = src/primal_dual_hybrid_gradient.jl   sp.jl  = src/saddle_point.jl
 *   isu.jl  = src/iteration_stats_utils.jl         tr.jl  = src/trust_region_utils.jl
 *   term.jl = src/termination.jl                   pre.jl = src/preprocess.jl
 *
 * Third-party arithmetic the reference delegates to the Julia stdlib
 * (SparseArrays `*`, LinearAlgebra norm/dot, Statistics.median) is restated
 * from its published algorithm:
 *   A*x   : column scatter  y[rowval[k]] += nzval[k]*x[j], j ascending
 *   A'*y  : per-column gather dot, k ascending
 *   norm(v,2) = sqrt(sum v_i^2), dot = sum x_i*y_i, left to right
 *   median: middle element, or a/2 + b/2 of the two middle elements
 * Parity with Julia is therefore pinned by the reference's known-answer tests
 * (answers), not by bit patterns of BLAS sums.
 */
#include "folp_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

/* ------------------------------------------------------------------------ */
/* small helpers                                                             */
/* ------------------------------------------------------------------------ */

static double now_sec(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

static double* dalloc(int64_t n) {
  double* p = (double*)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
  if (!p) { fprintf(stderr, "folp_oracle: out of memory\n"); abort(); }
  return p;
}
static double* dzeros(int64_t n) {
  double* p = (double*)calloc((size_t)(n > 0 ? n : 1), sizeof(double));
  if (!p) { fprintf(stderr, "folp_oracle: out of memory\n"); abort(); }
  return p;
}
static double* dcopy(const double* src, int64_t n) {
  double* p = dalloc(n);
  if (n > 0) memcpy(p, src, sizeof(double) * (size_t)n);
  return p;
}
static int64_t* ialloc(int64_t n) {
  int64_t* p = (int64_t*)malloc(sizeof(int64_t) * (size_t)(n > 0 ? n : 1));
  if (!p) { fprintf(stderr, "folp_oracle: out of memory\n"); abort(); }
  return p;
}

/* Julia's max/min on Float64 (NaN-propagating, -0.0 < +0.0). */
static double jl_max(double a, double b) {
  if (a != a) return a;
  if (b != b) return b;
  if (a > b) return a;
  if (b > a) return b;
  return signbit(a) ? b : a;
}
static double jl_min(double a, double b) {
  if (a != a) return a;
  if (b != b) return b;
  if (a < b) return a;
  if (b < a) return b;
  return signbit(a) ? a : b;
}
/* Base.clamp(x, lo, hi) = ifelse(x > hi, hi, ifelse(x < lo, lo, x)) */
static double jl_clamp(double x, double lo, double hi) {
  return x > hi ? hi : (x < lo ? lo : x);
}

static double norm2(const double* v, int64_t n) {
  double s = 0.0;
  for (int64_t i = 0; i < n; ++i) s += v[i] * v[i];
  return sqrt(s);
}
static double norminf(const double* v, int64_t n) {
  double s = 0.0;
  for (int64_t i = 0; i < n; ++i) {
    double a = fabs(v[i]);
    if (a != a) return a;
    if (a > s) s = a;
  }
  return s;
}
static double dot(const double* a, const double* b, int64_t n) {
  double s = 0.0;
  for (int64_t i = 0; i < n; ++i) s += a[i] * b[i];
  return s;
}

/* weighted_norm, sp.jl:120-129 */
static double weighted_norm(const double* v, const double* w, int64_t n) {
  double sum = 0.0;
  for (int64_t i = 0; i < n; ++i) sum += w[i] * v[i] * v[i];
  return sqrt(sum);
}

/* ------------------------------------------------------------------------ */
/* CSC matrix (SparseMatrixCSC{Float64,Int64}, 0-based here)                 */
/* ------------------------------------------------------------------------ */
typedef struct {
  int64_t m, n, nnz;
  int64_t* colptr; /* n+1 */
  int64_t* rowval; /* nnz */
  double* nzval;   /* nnz */
  int owns_pattern;
  /* CSR view (row pointers, column of every entry, its CSC position), built on first use by the
   * threaded A * x below; NULL otherwise */
  int64_t* t_rowptr;
  int64_t* t_col;
  int64_t* t_pos;
} csc_t;

/* Optional "all cores" timing mode (oracle_set_threads > 1; bench.py's stronger CPU baseline,
 * SURVEY.md section 8d). The reference's kernels are serial and so is the default. The threaded
 * products are bit-identical to the serial ones: A' * y is one independent dot product per
 * column, and A * x is evaluated row by row through a CSR view in ascending column order -- the
 * order in which the serial column scatter adds into each out[i]. */
static int g_threads = 1;
void oracle_set_threads(int threads) { g_threads = threads > 1 ? threads : 1; }
int oracle_get_threads(void) { return g_threads; }
int oracle_openmp_enabled(void) {
#ifdef _OPENMP
  return 1;
#else
  return 0;
#endif
}

static void csc_build_csr_view(csc_t* A) {
  int64_t* rp = (int64_t*)calloc((size_t)A->m + 1, sizeof(int64_t));
  int64_t* col = (int64_t*)malloc(sizeof(int64_t) * (size_t)(A->nnz > 0 ? A->nnz : 1));
  int64_t* pos = (int64_t*)malloc(sizeof(int64_t) * (size_t)(A->nnz > 0 ? A->nnz : 1));
  for (int64_t k = 0; k < A->nnz; ++k) rp[A->rowval[k] + 1] += 1;
  for (int64_t i = 0; i < A->m; ++i) rp[i + 1] += rp[i];
  int64_t* cur = (int64_t*)malloc(sizeof(int64_t) * (size_t)(A->m > 0 ? A->m : 1));
  for (int64_t i = 0; i < A->m; ++i) cur[i] = rp[i];
  for (int64_t j = 0; j < A->n; ++j)
    for (int64_t k = A->colptr[j]; k < A->colptr[j + 1]; ++k) {
      int64_t at = cur[A->rowval[k]]++;
      col[at] = j;
      pos[at] = k;
    }
  free(cur);
  A->t_rowptr = rp; A->t_col = col; A->t_pos = pos;
}

/* SparseArrays: mul!(C, A, B) column scatter */
static void csc_mul(const csc_t* A, const double* x, double* out) {
  if (g_threads > 1 && A->nnz > 0) {
    if (!A->t_rowptr) csc_build_csr_view((csc_t*)A); /* cache; the handle is used by one thread */
    const int64_t* rp = A->t_rowptr;
#pragma omp parallel for num_threads(g_threads) schedule(static)
    for (int64_t i = 0; i < A->m; ++i) {
      double tmp = 0.0;
      for (int64_t k = rp[i]; k < rp[i + 1]; ++k) tmp += A->nzval[A->t_pos[k]] * x[A->t_col[k]];
      out[i] = tmp;
    }
    return;
  }
  for (int64_t i = 0; i < A->m; ++i) out[i] = 0.0;
  for (int64_t j = 0; j < A->n; ++j) {
    double xj = x[j];
    for (int64_t k = A->colptr[j]; k < A->colptr[j + 1]; ++k)
      out[A->rowval[k]] += A->nzval[k] * xj;
  }
}
/* SparseArrays: mul!(C, adjoint(A), B) per-column gather */
static void csc_tmul(const csc_t* A, const double* y, double* out) {
#pragma omp parallel for num_threads(g_threads) schedule(static) if (g_threads > 1)
  for (int64_t j = 0; j < A->n; ++j) {
    double tmp = 0.0;
    for (int64_t k = A->colptr[j]; k < A->colptr[j + 1]; ++k)
      tmp += A->nzval[k] * y[A->rowval[k]];
    out[j] = tmp;
  }
}

static void csc_free(csc_t* A) {
  if (A->owns_pattern) { free(A->colptr); free(A->rowval); }
  free(A->t_rowptr); free(A->t_col); free(A->t_pos);
  A->t_rowptr = A->t_col = A->t_pos = NULL;
  free(A->nzval);
  memset(A, 0, sizeof(*A));
}

/* ------------------------------------------------------------------------ */
/* QuadraticProgrammingProblem, quadratic_programming.jl:34-76               */
/* ------------------------------------------------------------------------ */
typedef struct {
  int64_t n, m, neq;
  double *l, *u, *c, *b;
  double c0;
  csc_t A;
  csc_t Q; /* nnz == 0 for an LP (is_linear_programming_problem) */
} qp_t;

static int qp_is_lp(const qp_t* p) {
  /* iszero(objective_matrix) */
  for (int64_t k = 0; k < p->Q.nnz; ++k)
    if (p->Q.nzval[k] != 0.0) return 0;
  return 1;
}

static void build_pattern(int64_t ncols, int64_t nnz, int base,
                          const int64_t* colptr, const int64_t* rowval,
                          csc_t* out) {
  out->colptr = ialloc(ncols + 1);
  out->rowval = ialloc(nnz);
  if (colptr) {
    for (int64_t j = 0; j <= ncols; ++j) out->colptr[j] = colptr[j] - base;
  } else {
    for (int64_t j = 0; j <= ncols; ++j) out->colptr[j] = 0;
  }
  for (int64_t k = 0; k < nnz; ++k) out->rowval[k] = rowval[k] - base;
  out->owns_pattern = 1;
}

/* which = 0: the scaled problem of *fp; which = 1: the original problem. */
static void qp_from_folp(const folp_problem* fp, int which, qp_t* out) {
  int64_t n = fp->num_variables, m = fp->num_constraints;
  int64_t nnz = fp->num_nonzeros;
  memset(out, 0, sizeof(*out));
  out->n = n; out->m = m; out->neq = fp->num_equalities;
  out->c0 = fp->objective_constant;
  out->A.m = m; out->A.n = n; out->A.nnz = nnz;
  build_pattern(n, nnz, fp->index_base, fp->colptr, fp->rowval, &out->A);
  out->Q.m = n; out->Q.n = n; out->Q.nnz = fp->q_num_nonzeros;
  build_pattern(n, fp->q_num_nonzeros, fp->index_base,
                fp->q_num_nonzeros > 0 ? fp->q_colptr : NULL, fp->q_rowval,
                &out->Q);
  if (which == 0) {
    out->l = dcopy(fp->variable_lower_bound, n);
    out->u = dcopy(fp->variable_upper_bound, n);
    out->c = dcopy(fp->objective_vector, n);
    out->b = dcopy(fp->right_hand_side, m);
    out->A.nzval = dcopy(fp->nzval, nnz);
    out->Q.nzval = dcopy(fp->q_nzval, fp->q_num_nonzeros);
  } else {
    const double* D = fp->variable_rescaling;
    const double* E = fp->constraint_rescaling;
    out->l = dcopy(fp->orig_variable_lower_bound ? fp->orig_variable_lower_bound
                                                 : fp->variable_lower_bound, n);
    out->u = dcopy(fp->orig_variable_upper_bound ? fp->orig_variable_upper_bound
                                                 : fp->variable_upper_bound, n);
    out->c = dcopy(fp->orig_objective_vector ? fp->orig_objective_vector
                                             : fp->objective_vector, n);
    out->b = dcopy(fp->orig_right_hand_side ? fp->orig_right_hand_side
                                            : fp->right_hand_side, m);
    if (fp->orig_nzval) {
      out->A.nzval = dcopy(fp->orig_nzval, nnz);
    } else {
      /* fallback when the caller has no original values: undo
       * constraint_matrix = E^-1 A D^-1 (pre.jl:569-571) */
      out->A.nzval = dcopy(fp->nzval, nnz);
      for (int64_t j = 0; j < n; ++j)
        for (int64_t k = out->A.colptr[j]; k < out->A.colptr[j + 1]; ++k)
          out->A.nzval[k] = out->A.nzval[k] * (E ? E[out->A.rowval[k]] : 1.0) *
                            (D ? D[j] : 1.0);
    }
    if (fp->q_orig_nzval) {
      out->Q.nzval = dcopy(fp->q_orig_nzval, fp->q_num_nonzeros);
    } else {
      out->Q.nzval = dcopy(fp->q_nzval, fp->q_num_nonzeros);
      for (int64_t j = 0; j < n; ++j)
        for (int64_t k = out->Q.colptr[j]; k < out->Q.colptr[j + 1]; ++k)
          out->Q.nzval[k] = out->Q.nzval[k] * (D ? D[out->Q.rowval[k]] : 1.0) *
                            (D ? D[j] : 1.0);
    }
  }
}

static void qp_free(qp_t* p) {
  free(p->l); free(p->u); free(p->c); free(p->b);
  csc_free(&p->A); csc_free(&p->Q);
}

/* ------------------------------------------------------------------------ */
/* projections, sp.jl:82-117                                                 */
/* ------------------------------------------------------------------------ */
static void project_primal(double* x, const qp_t* p) {
  for (int64_t i = 0; i < p->n; ++i)
    x[i] = jl_min(p->u[i], jl_max(p->l[i], x[i]));
}
static void project_dual(double* y, const qp_t* p) {
  for (int64_t i = p->neq; i < p->m; ++i) y[i] = jl_max(y[i], 0.0);
}

/* compute_primal_gradient_from_dual_product, sp.jl:1093-1100:
 * objective_matrix * x .+ objective_vector .- dual_product */
static void primal_gradient_from_dual_product(const qp_t* p, const double* x,
                                              const double* dual_product,
                                              double* out) {
  double* qx = dalloc(p->n);
  csc_mul(&p->Q, x, qx);
  for (int64_t j = 0; j < p->n; ++j) out[j] = (qx[j] + p->c[j]) - dual_product[j];
  free(qx);
}
/* compute_primal_gradient, sp.jl:1081-1091 */
static void primal_gradient(const qp_t* p, const double* x, const double* y,
                            double* out) {
  double* aty = dalloc(p->n);
  csc_tmul(&p->A, y, aty);
  primal_gradient_from_dual_product(p, x, aty, out);
  free(aty);
}
/* compute_dual_gradient, sp.jl:1102-1107 */
static void dual_gradient(const qp_t* p, const double* x, double* out) {
  double* ax = dalloc(p->m);
  csc_mul(&p->A, x, ax);
  for (int64_t i = 0; i < p->m; ++i) out[i] = p->b[i] - ax[i];
  free(ax);
}
/* compute_lagrangian_value, sp.jl:1109-1120 */
static double lagrangian_value(const qp_t* p, const double* x, const double* y) {
  double* qx = dalloc(p->n);
  double* aty = dalloc(p->n);
  csc_mul(&p->Q, x, qx);
  csc_tmul(&p->A, y, aty);
  double v = 0.5 * dot(x, qx, p->n) + dot(x, p->c, p->n) - dot(x, aty, p->n) +
             dot(y, p->b, p->m) + p->c0;
  free(qx); free(aty);
  return v;
}

/* ------------------------------------------------------------------------ */
/* trust region, tr.jl:68-224                                                */
/* ------------------------------------------------------------------------ */
static int cmp_double(const void* a, const void* b) {
  double x = *(const double*)a, y = *(const double*)b;
  return (x > y) - (x < y);
}
/* quickselect: k-th smallest (0-based) of v[0..n), v is permuted */
static double select_kth(double* v, int64_t n, int64_t k) {
  int64_t lo = 0, hi = n - 1;
  while (lo < hi) {
    double pivot = v[lo + (hi - lo) / 2];
    int64_t i = lo, j = hi;
    while (i <= j) {
      while (v[i] < pivot) ++i;
      while (v[j] > pivot) --j;
      if (i <= j) { double t = v[i]; v[i] = v[j]; v[j] = t; ++i; --j; }
    }
    if (k <= j) hi = j;
    else if (k >= i) lo = i;
    else return v[k];
  }
  return v[k];
}
/* Statistics.median: middle(x, y) = x/2 + y/2 for an even count */
static double median_of(double* scratch, int64_t n) {
  if (n <= 64) {
    qsort(scratch, (size_t)n, sizeof(double), cmp_double);
    if (n & 1) return scratch[n / 2];
    return scratch[n / 2 - 1] / 2 + scratch[n / 2] / 2;
  }
  if (n & 1) return select_kth(scratch, n, n / 2);
  double hi = select_kth(scratch, n, n / 2);
  /* after selection everything left of n/2 is <= hi; the lower middle is the
   * maximum of that part */
  double lo = scratch[0];
  for (int64_t i = 1; i < n / 2; ++i)
    if (scratch[i] > lo) lo = scratch[i];
  return lo / 2 + hi / 2;
}

/* approximately_solve_bound_constrained_trust_region, tr.jl:194-224 */
static double tr_approx(int64_t len, const double* center, const double* obj,
                        const double* lb, const double* ub, const double* w,
                        double radius, double* solution) {
  double* direction = dzeros(len);
  for (int64_t i = 0; i < len; ++i) {
    if (center[i] >= ub[i] && obj[i] <= 0) continue;
    if (center[i] <= lb[i] && obj[i] >= 0) continue;
    direction[i] = -obj[i] / w[i];
  }
  double direction_norm = weighted_norm(direction, w, len);
  if (direction_norm > 0.0) {
    double f = radius / direction_norm;
    for (int64_t i = 0; i < len; ++i) direction[i] *= f;
  }
  for (int64_t i = 0; i < len; ++i) solution[i] = center[i] + direction[i];
  double value = dot(obj, direction, len);
  free(direction);
  return value;
}

/* solve_bound_constrained_trust_region, tr.jl:68-192 */
static double tr_solve(int64_t len, const double* center, const double* obj,
                       const double* lb, const double* ub, const double* w,
                       double radius, int approx, double* solution) {
  if (approx) return tr_approx(len, center, obj, lb, ub, w, radius, solution);
  /* :88-91 */
  if (radius == 0.0 || norm2(obj, len) == 0.0) {
    memcpy(solution, center, sizeof(double) * (size_t)len);
    return 0.0;
  }
  double* direction = dzeros(len);
  double* threshold = dzeros(len);
  for (int64_t i = 0; i < len; ++i) { /* :95-117 */
    if (center[i] >= ub[i] && obj[i] <= 0) continue;
    if (center[i] <= lb[i] && obj[i] >= 0) continue;
    direction[i] = -obj[i] / w[i];
    if (direction[i] > 0) threshold[i] = (ub[i] - center[i]) / direction[i];
    else if (direction[i] < 0) threshold[i] = (lb[i] - center[i]) / direction[i];
    else threshold[i] = 0.0;
  }
  double low_radius_sq = 0.0, high_radius_sq = 0.0;
  int64_t* indices = ialloc(len);
  int64_t nidx = 0;
  { /* :126-132 */
    double s = 0.0;
    for (int64_t i = 0; i < len; ++i)
      if (isinf(threshold[i])) s += w[i] * direction[i] * direction[i];
    double wn = sqrt(s);
    high_radius_sq += wn * wn;
    for (int64_t i = 0; i < len; ++i)
      if (isfinite(threshold[i])) indices[nidx++] = i;
  }
  double* scratch = dalloc(nidx);
  while (nidx > 0) { /* :134-171 */
    for (int64_t k = 0; k < nidx; ++k) scratch[k] = threshold[indices[k]];
    double test_threshold = median_of(scratch, nidx);
    double s = 0.0;
    for (int64_t k = 0; k < nidx; ++k) {
      int64_t i = indices[k];
      double tp = jl_clamp(center[i] + test_threshold * direction[i], lb[i], ub[i]);
      double d = tp - center[i];
      s += w[i] * d * d;
    }
    double test_radius = sqrt(s);
    if (low_radius_sq + test_radius * test_radius +
            test_threshold * test_threshold * high_radius_sq >=
        radius * radius) {
      double sd = 0.0;
      int64_t keep = 0;
      for (int64_t k = 0; k < nidx; ++k) {
        int64_t i = indices[k];
        if (threshold[i] >= test_threshold) sd += w[i] * direction[i] * direction[i];
        else indices[keep++] = i;
      }
      double wn = sqrt(sd);
      high_radius_sq += wn * wn;
      nidx = keep;
    } else {
      double sd = 0.0;
      int64_t keep = 0;
      for (int64_t k = 0; k < nidx; ++k) {
        int64_t i = indices[k];
        if (threshold[i] <= test_threshold) {
          double dp = jl_clamp(center[i] + test_threshold * direction[i], lb[i], ub[i]);
          double d = dp - center[i];
          sd += w[i] * d * d;
        }
        if (threshold[i] > test_threshold) indices[keep++] = i;
      }
      double wn = sqrt(sd);
      low_radius_sq += wn * wn;
      nidx = keep;
    }
  }
  double target_threshold; /* :175-181 */
  if (high_radius_sq <= 0.0) {
    target_threshold = threshold[0];
    for (int64_t i = 1; i < len; ++i)
      target_threshold = jl_max(target_threshold, threshold[i]);
  } else {
    target_threshold = sqrt((radius * radius - low_radius_sq) / high_radius_sq);
  }
  double value = 0.0; /* :182-191 */
  for (int64_t i = 0; i < len; ++i) {
    solution[i] = jl_clamp(center[i] + target_threshold * direction[i], lb[i], ub[i]);
    value += obj[i] * (solution[i] - center[i]);
  }
  free(direction); free(threshold); free(indices); free(scratch);
  return value;
}

/* OptimalObjectiveBoundResult, tr.jl:226-238 (vectors returned on request) */
typedef struct {
  double lagrangian_value, lower_bound_value, upper_bound_value;
} bound_result;
static double get_gap(const bound_result* r) {
  return r->upper_bound_value - r->lower_bound_value;
}

/* bound_optimal_objective, tr.jl:271-360 */
static bound_result bound_optimal_objective(const qp_t* p, const double* x,
                                            const double* y, const double* wp,
                                            const double* wd, double radius,
                                            int norm_kind, int approx,
                                            double* x_tr, double* y_tr) {
  int64_t n = p->n, m = p->m;
  bound_result r;
  double* g = dalloc(n);
  primal_gradient(p, x, y, g);                 /* :282-283 */
  r.lagrangian_value = lagrangian_value(p, x, y); /* :285-286 */
  double* dlb = dalloc(m);
  double* dub = dalloc(m);
  for (int64_t i = 0; i < m; ++i) { dlb[i] = -INFINITY; dub[i] = INFINITY; }
  for (int64_t i = p->neq; i < m; ++i) dlb[i] = 0.0;
  double* gy = dalloc(m);
  dual_gradient(p, x, gy);                     /* :291 */
  if (norm_kind == 0) {                        /* MAX_NORM :293-326 */
    double* xs = dalloc(n);
    double* ys = dalloc(m);
    double pv = tr_solve(n, x, g, p->l, p->u, wp, radius, approx, xs);
    double* neg = dalloc(m);
    for (int64_t i = 0; i < m; ++i) neg[i] = -gy[i];
    double dv = tr_solve(m, y, neg, dlb, dub, wd, radius, approx, ys);
    r.lower_bound_value = r.lagrangian_value + pv;
    r.upper_bound_value = r.lagrangian_value - dv;
    if (x_tr) memcpy(x_tr, xs, sizeof(double) * (size_t)n);
    if (y_tr) memcpy(y_tr, ys, sizeof(double) * (size_t)m);
    free(xs); free(ys); free(neg);
  } else {                                     /* EUCLIDEAN_NORM :327-356 */
    int64_t len = n + m;
    double* z = dalloc(len); double* zg = dalloc(len);
    double* zl = dalloc(len); double* zu = dalloc(len);
    double* zw = dalloc(len); double* zs = dalloc(len);
    for (int64_t j = 0; j < n; ++j) {
      z[j] = x[j]; zg[j] = g[j]; zl[j] = p->l[j]; zu[j] = p->u[j]; zw[j] = wp[j];
    }
    for (int64_t i = 0; i < m; ++i) {
      z[n + i] = y[i]; zg[n + i] = -gy[i]; zl[n + i] = dlb[i]; zu[n + i] = dub[i];
      zw[n + i] = wd[i];
    }
    tr_solve(len, z, zg, zl, zu, zw, radius, approx, zs);
    double lo = 0.0, up = 0.0;
    for (int64_t j = 0; j < n; ++j) lo += (zs[j] - x[j]) * g[j];
    for (int64_t i = 0; i < m; ++i) up += (zs[n + i] - y[i]) * gy[i];
    r.lower_bound_value = r.lagrangian_value + lo;
    r.upper_bound_value = r.lagrangian_value + up;
    if (x_tr) memcpy(x_tr, zs, sizeof(double) * (size_t)n);
    if (y_tr) memcpy(y_tr, zs + n, sizeof(double) * (size_t)m);
    free(z); free(zg); free(zl); free(zu); free(zw); free(zs);
  }
  free(g); free(dlb); free(dub); free(gy);
  return r;
}

/* ------------------------------------------------------------------------ */
/* iteration stats, isu.jl                                                   */
/* ------------------------------------------------------------------------ */

/* compute_primal_residual, isu.jl:30-63; out has m + 2n entries */
static void primal_residual(const qp_t* p, const double* x, double* out) {
  int64_t m = p->m, n = p->n;
  double* act = dalloc(m);
  csc_mul(&p->A, x, act);
  for (int64_t i = 0; i < p->neq; ++i) out[i] = p->b[i] - act[i];
  for (int64_t i = p->neq; i < m; ++i) out[i] = jl_max(p->b[i] - act[i], 0.0);
  for (int64_t j = 0; j < n; ++j) out[m + j] = jl_max(p->l[j] - x[j], 0.0);
  for (int64_t j = 0; j < n; ++j) out[m + n + j] = jl_max(x[j] - p->u[j], 0.0);
  free(act);
}
/* primal_obj, isu.jl:67-74 */
static double primal_obj(const qp_t* p, const double* x) {
  double* qx = dalloc(p->n);
  /* (x' * Q) * x : row-vector times sparse = (Q' x)' */
  csc_tmul(&p->Q, x, qx);
  double v = p->c0 + dot(p->c, x, p->n) + 0.5 * dot(qx, x, p->n);
  free(qx);
  return v;
}
/* reduced_costs_dual_objective_contribution, isu.jl:93-117 */
static double rc_dual_objective_contribution(const double* l, const double* u,
                                             const double* rc, int64_t n) {
  double s = 0.0;
  for (int64_t i = 0; i < n; ++i) {
    double bound;
    if (rc[i] == 0.0) continue;
    else if (rc[i] > 0.0) bound = l[i];
    else bound = u[i];
    if (!isfinite(bound)) return -INFINITY;
    s += bound * rc[i];
  }
  return s;
}
/* compute_reduced_costs_from_primal_gradient, isu.jl:128-147 */
static void reduced_costs_from_gradient(const double* l, const double* u,
                                        const double* g, int64_t n, double* rc) {
  for (int64_t i = 0; i < n; ++i) {
    double bound = g[i] > 0.0 ? l[i] : u[i];
    rc[i] = isfinite(bound) ? g[i] : 0.0;
  }
}
/* compute_dual_stats, isu.jl:157-197. dres: (m-neq)+n, rc: n */
static double dual_stats(const qp_t* p, const double* x, const double* y,
                         double* dres, double* rc) {
  int64_t n = p->n, m = p->m, nin = m - p->neq;
  double* qx = dalloc(n);
  double* g = dalloc(n);
  csc_mul(&p->Q, x, qx);
  primal_gradient(p, x, y, g);
  reduced_costs_from_gradient(p->l, p->u, g, n, rc);
  for (int64_t i = 0; i < nin; ++i) dres[i] = jl_max(-y[p->neq + i], 0.0);
  for (int64_t j = 0; j < n; ++j) dres[nin + j] = g[j] - rc[j];
  double base = dot(p->b, y, m) + p->c0 - 0.5 * dot(qx, x, n);
  double dobj = base + rc_dual_objective_contribution(p->l, p->u, rc, n);
  free(qx); free(g);
  return dobj;
}

/* compute_convergence_information, isu.jl:228-280 */
static void convergence_information(const qp_t* p, const double cache[4],
                                    const double* x, const double* y,
                                    double eps_ratio, int candidate_type,
                                    folp_eval* e) {
  int64_t n = p->n, m = p->m, nin = m - p->neq;
  double* pres = dalloc(m + 2 * n);
  primal_residual(p, x, pres);
  e->primal_objective = primal_obj(p, x);
  e->l_inf_primal_residual = norminf(pres, m + 2 * n);
  e->l2_primal_residual = norm2(pres, m + 2 * n);
  e->relative_l_inf_primal_residual = e->l_inf_primal_residual / (eps_ratio + cache[1]);
  e->relative_l2_primal_residual = e->l2_primal_residual / (eps_ratio + cache[3]);
  e->l_inf_primal_variable = norminf(x, n);
  e->l2_primal_variable = norm2(x, n);
  double* dres = dalloc(nin + n);
  double* rc = dalloc(n);
  e->dual_objective = dual_stats(p, x, y, dres, rc);
  e->l_inf_dual_residual = norminf(dres, nin + n);
  e->l2_dual_residual = norm2(dres, nin + n);
  e->relative_l_inf_dual_residual = e->l_inf_dual_residual / (eps_ratio + cache[0]);
  e->relative_l2_dual_residual = e->l2_dual_residual / (eps_ratio + cache[2]);
  e->l_inf_dual_variable = norminf(y, m);
  e->l2_dual_variable = norm2(y, m);
  /* corrected_dual_obj, isu.jl:203-212 */
  e->corrected_dual_objective =
      (e->l_inf_dual_residual == 0.0) ? e->dual_objective : -INFINITY;
  double gap = fabs(e->primal_objective - e->dual_objective);
  double abs_obj = fabs(e->primal_objective) + fabs(e->dual_objective);
  e->relative_optimality_gap = gap / (eps_ratio + abs_obj);
  e->candidate_type = candidate_type;
  free(pres); free(dres); free(rc);
}

/* compute_infeasibility_information, isu.jl:287-349 */
static void infeasibility_information(const qp_t* p, const double* primal_ray_in,
                                      const double* dual_ray, folp_eval* e) {
  int64_t n = p->n, m = p->m, nin = m - p->neq;
  double* xr = dcopy(primal_ray_in, n);
  double inf_norm = norminf(xr, n);
  if (inf_norm != 0.0)
    for (int64_t j = 0; j < n; ++j) xr[j] /= inf_norm;
  if (g_threads > 1) { /* the copies below must share the CSR views, not build (and leak) their own */
    if (!p->A.t_rowptr && p->A.nnz > 0) csc_build_csr_view((csc_t*)&p->A);
    if (!p->Q.t_rowptr && p->Q.nnz > 0) csc_build_csr_view((csc_t*)&p->Q);
  }
  /* homogeneous primal :301-309 */
  qp_t hp = *p;
  hp.l = dalloc(n); hp.u = dalloc(n); hp.b = dzeros(m);
  for (int64_t j = 0; j < n; ++j) {
    hp.l[j] = isfinite(p->l[j]) ? 0.0 : -INFINITY;
    hp.u[j] = isfinite(p->u[j]) ? 0.0 : INFINITY;
  }
  double* hres = dalloc(m + 2 * n);
  primal_residual(&hp, xr, hres);
  e->max_primal_ray_infeasibility = norminf(hres, m + 2 * n);
  e->primal_ray_linear_objective = dot(p->c, xr, n);
  double* qx = dalloc(n);
  csc_mul(&p->Q, xr, qx);
  e->primal_ray_quadratic_norm = norminf(qx, n);
  free(qx); free(hres); free(hp.l); free(hp.u); free(hp.b);
  /* homogeneous dual :319-330: LP with c = 0, c0 = 0, Q = 0 */
  qp_t hd = *p;
  hd.c = dzeros(n); hd.c0 = 0.0;
  hd.Q.nnz = 0;
  int64_t* zero_colptr = (int64_t*)calloc((size_t)n + 1, sizeof(int64_t));
  hd.Q.colptr = zero_colptr;
  hd.Q.t_rowptr = hd.Q.t_col = hd.Q.t_pos = NULL; /* the copy has another pattern than p->Q */
  double* dres = dalloc(nin + n);
  double* rc = dalloc(n);
  double dobj = dual_stats(&hd, xr, dual_ray, dres, rc);
  double scaling_factor = jl_max(norminf(dual_ray, m), norminf(rc, n));
  if (scaling_factor != 0.0) {
    e->max_dual_ray_infeasibility = norminf(dres, nin + n) / scaling_factor;
    e->dual_ray_objective = dobj / scaling_factor;
  } else {
    e->max_dual_ray_infeasibility = 0.0;
    e->dual_ray_objective = 0.0;
  }
  free(hd.c); free(zero_colptr); free(dres); free(rc); free(xr);
}

/* ------------------------------------------------------------------------ */
/* termination, term.jl:163-273                                              */
/* ------------------------------------------------------------------------ */
static int optimality_criteria_met(int norm, double abs_tol, double rel_tol,
                                   const folp_eval* ci, const double cache[4]) {
  double abs_obj = fabs(ci->primal_objective) + fabs(ci->dual_objective);
  double gap = fabs(ci->primal_objective - ci->dual_objective);
  double perr, pbase, derr, dbase;
  if (norm == FOLP_L_INF) {
    perr = ci->l_inf_primal_residual; pbase = cache[1];
    derr = ci->l_inf_dual_residual; dbase = cache[0];
  } else {
    perr = ci->l2_primal_residual; pbase = cache[3];
    derr = ci->l2_dual_residual; dbase = cache[2];
  }
  return derr < abs_tol + rel_tol * dbase && perr < abs_tol + rel_tol * pbase &&
         gap < abs_tol + rel_tol * abs_obj;
}
static int primal_infeasibility_criteria_met(double eps, const folp_eval* ii) {
  if (ii->dual_ray_objective <= 0.0) return 0;
  return ii->max_dual_ray_infeasibility / ii->dual_ray_objective <= eps;
}
static int dual_infeasibility_criteria_met(double eps, const folp_eval* ii) {
  if (ii->primal_ray_linear_objective >= 0.0) return 0;
  return ii->max_primal_ray_infeasibility / (-ii->primal_ray_linear_objective) <= eps &&
         ii->primal_ray_quadratic_norm / (-ii->primal_ray_linear_objective) <= eps;
}
static int check_termination(const folp_params* c, const double cache[4],
                             const folp_eval* s) {
  if (optimality_criteria_met(c->optimality_norm, c->eps_optimal_absolute,
                              c->eps_optimal_relative, s, cache))
    return FOLP_TERMINATION_REASON_OPTIMAL;
  if (primal_infeasibility_criteria_met(c->eps_primal_infeasible, s))
    return FOLP_TERMINATION_REASON_PRIMAL_INFEASIBLE;
  if (dual_infeasibility_criteria_met(c->eps_dual_infeasible, s))
    return FOLP_TERMINATION_REASON_DUAL_INFEASIBLE;
  if (s->iteration_number >= c->iteration_limit)
    return FOLP_TERMINATION_REASON_ITERATION_LIMIT;
  else if (s->cumulative_kkt_matrix_passes >= c->kkt_matrix_pass_limit)
    return FOLP_TERMINATION_REASON_KKT_MATRIX_PASS_LIMIT;
  else if (s->cumulative_time_sec >= c->time_sec_limit)
    return FOLP_TERMINATION_REASON_TIME_LIMIT;
  return 0;
}

/* ------------------------------------------------------------------------ */
/* solver state, pdhg.jl:205-258; averages sp.jl:215-301; RestartInfo :158   */
/* ------------------------------------------------------------------------ */
struct oracle_handle {
  qp_t P;  /* scaled problem  */
  qp_t O;  /* original problem */
  double *D, *E; /* variable_rescaling, constraint_rescaling */
  double cache[4];
  folp_params prm;
  int p_is_lp;
  /* PdhgSolverState */
  double *x, *y, *dx, *dy, *aty;
  double *sum_x, *sum_y;
  int64_t count_x, count_y;
  double w_x, w_y;
  double step_size, primal_weight;
  int numerical_error;
  double kkt_passes;
  int64_t total_number_iterations;
  double ratio_step_sizes;
  /* RestartInfo */
  double *last_x, *last_y;
  int has_last_gap;
  double last_gap;
  int64_t last_restart_length;
  double pd_last, dd_last, gap_reduction_ratio_last_trial;
  /* loop bookkeeping */
  int64_t iteration; /* the reference's `iteration` */
  int need_step, terminated;
  double start_time, basic_time;
  double last_interaction, last_movement;
  folp_eval last_eval;
  /* attempt-wise stepping (oracle_debug_attempts only) */
  int in_retry;
  double entry_weight;
};

/* add_to_*_solution_weighted_average, sp.jl:252-294 */
static void add_primal_avg(oracle_handle* h, const double* x, double w) {
  for (int64_t j = 0; j < h->P.n; ++j) h->sum_x[j] += x[j] * w;
  h->count_x += 1;
  h->w_x += w;
}
static void add_dual_avg(oracle_handle* h, const double* y, double w) {
  for (int64_t i = 0; i < h->P.m; ++i) h->sum_y[i] += y[i] * w;
  h->count_y += 1;
  h->w_y += w;
}

/* compute_next_primal_solution, pdhg.jl:442-470 */
static double* next_primal_solution(const oracle_handle* h, double step_size) {
  const qp_t* p = &h->P;
  double* g = dalloc(p->n);
  primal_gradient_from_dual_product(p, h->x, h->aty, g);
  double* next = dalloc(p->n);
  double f = step_size / h->primal_weight;
  for (int64_t j = 0; j < p->n; ++j) next[j] = h->x[j] - f * g[j];
  project_primal(next, p);
  free(g);
  return next;
}
/* compute_next_dual_solution, pdhg.jl:472-494 */
static void next_dual_solution(const oracle_handle* h, const double* next_primal,
                               double step_size, double extrapolation,
                               double** next_dual_out, double** next_aty_out) {
  const qp_t* p = &h->P;
  double* xbar = dalloc(p->n);
  for (int64_t j = 0; j < p->n; ++j)
    xbar[j] = next_primal[j] + extrapolation * (next_primal[j] - h->x[j]);
  double* gy = dalloc(p->m);
  dual_gradient(p, xbar, gy);
  double* next = dalloc(p->m);
  double f = h->primal_weight * step_size;
  for (int64_t i = 0; i < p->m; ++i) next[i] = h->y[i] + f * gy[i];
  project_dual(next, p);
  double* naty = dalloc(p->n);
  csc_tmul(&p->A, next, naty);
  free(xbar); free(gy);
  *next_dual_out = next; *next_aty_out = naty;
}
/* update_solution_in_solver_state, pdhg.jl:500-519 (takes ownership) */
static void update_solution(oracle_handle* h, double* np, double* nd, double* naty,
                            double weight) {
  for (int64_t j = 0; j < h->P.n; ++j) h->dx[j] = np[j] - h->x[j];
  for (int64_t i = 0; i < h->P.m; ++i) h->dy[i] = nd[i] - h->y[i];
  free(h->x); free(h->y); free(h->aty);
  h->x = np; h->y = nd; h->aty = naty;
  /* weight = solver_state.step_size, i.e. the value at ENTRY to take_step
   * (:512; the field is only written back at :730/:644) */
  add_primal_avg(h, h->x, weight);
  add_dual_avg(h, h->y, weight);
}
/* compute_interaction_and_movement, pdhg.jl:527-549 */
static void interaction_and_movement(const oracle_handle* h, const double* np,
                                     const double* nd, const double* naty,
                                     double* interaction, double* movement) {
  const qp_t* p = &h->P;
  double* dxv = dalloc(p->n);
  double* dyv = dalloc(p->m);
  for (int64_t j = 0; j < p->n; ++j) dxv[j] = np[j] - h->x[j];
  for (int64_t i = 0; i < p->m; ++i) dyv[i] = nd[i] - h->y[i];
  double obj_inter = 0.0;
  if (!h->p_is_lp) {
    double* q = dalloc(p->n);
    csc_tmul(&p->Q, dxv, q); /* (dx' * Q) */
    obj_inter = 0.5 * dot(q, dxv, p->n);
    free(q);
  }
  double pd = 0.0;
  for (int64_t j = 0; j < p->n; ++j) pd += dxv[j] * (naty[j] - h->aty[j]);
  *interaction = fabs(pd) + fabs(obj_inter);
  double ndx = norm2(dxv, p->n), ndy = norm2(dyv, p->m);
  *movement = 0.5 * h->primal_weight * (ndx * ndx) +
              (0.5 / h->primal_weight) * (ndy * ndy);
  free(dxv); free(dyv);
}

/* take_step(::AdaptiveStepsizeParams), pdhg.jl:653-731.
 * max_attempts < 0: run to acceptance (the reference). Otherwise stop after
 * that many inner attempts (test hook); returns attempts used, and *done. */
static int64_t take_step_adaptive(oracle_handle* h, int64_t max_attempts, int* done_out) {
  double step_size = h->step_size;
  int done = 0;
  int64_t iter = 0;
  while (!done && (max_attempts < 0 || iter < max_attempts)) {
    iter += 1;
    h->total_number_iterations += 1;
    double* np = next_primal_solution(h, step_size);
    double *nd, *naty;
    next_dual_solution(h, np, step_size, 1.0, &nd, &naty);
    double interaction, movement;
    interaction_and_movement(h, np, nd, naty, &interaction, &movement);
    h->last_interaction = interaction; h->last_movement = movement;
    h->kkt_passes += 1;
    if (movement == 0.0) {
      h->numerical_error = 1;
      free(np); free(nd); free(naty);
      break;
    }
    double step_size_limit = interaction > 0 ? movement / interaction : INFINITY;
    if (step_size <= step_size_limit) {
      update_solution(h, np, nd, naty, h->step_size);
      done = 1;
    } else {
      free(np); free(nd); free(naty);
    }
    double k1 = (double)(h->total_number_iterations + 1);
    double first_term = (1 - pow(k1, -h->prm.reduction_exponent)) * step_size_limit;
    double second_term = (1 + pow(k1, -h->prm.growth_exponent)) * step_size;
    step_size = jl_min(first_term, second_term);
  }
  h->step_size = step_size;
  if (done_out) *done_out = done;
  return iter;
}

/* take_step(::MalitskyPockStepsizeParameters), pdhg.jl:555-647 */
static void take_step_malitsky_pock(oracle_handle* h) {
  double step_size = h->step_size;
  double ratio_step_sizes = h->ratio_step_sizes;
  int done = 0;
  int iter = 0;
  double* np = next_primal_solution(h, step_size);
  h->kkt_passes += 0.5;
  step_size = step_size +
              h->prm.interpolation_coefficient * (sqrt(1 + ratio_step_sizes) - 1) * step_size;
  const int max_iter = 60;
  while (!done && iter < max_iter) {
    iter += 1;
    h->total_number_iterations += 1;
    ratio_step_sizes = step_size / h->step_size;
    double *nd, *naty;
    next_dual_solution(h, np, step_size, ratio_step_sizes, &nd, &naty);
    double sdy = 0.0, sdp = 0.0;
    for (int64_t i = 0; i < h->P.m; ++i) { double d = nd[i] - h->y[i]; sdy += d * d; }
    for (int64_t j = 0; j < h->P.n; ++j) { double d = naty[j] - h->aty[j]; sdp += d * d; }
    h->kkt_passes += 0.5;
    if (step_size * sqrt(sdp) <= h->prm.breaking_factor * sqrt(sdy)) {
      if (h->count_x == 0) add_primal_avg(h, h->x, step_size * ratio_step_sizes);
      update_solution(h, np, nd, naty, h->step_size);
      np = NULL;
      done = 1;
    } else {
      step_size *= h->prm.downscaling_factor;
      free(nd); free(naty);
    }
  }
  if (iter == max_iter && !done) {
    h->numerical_error = 1;
    free(np);
    return;
  }
  h->step_size = step_size;
  h->ratio_step_sizes = ratio_step_sizes;
}

/* take_step(::ConstantStepsizeParams), pdhg.jl:737-767 */
static void take_step_constant(oracle_handle* h) {
  double* np = next_primal_solution(h, h->step_size);
  double *nd, *naty;
  next_dual_solution(h, np, h->step_size, 1.0, &nd, &naty);
  h->kkt_passes += 1;
  update_solution(h, np, nd, naty, h->step_size);
}

static void take_step(oracle_handle* h) {
  double t0 = now_sec();
  switch (h->prm.step_size_policy) {
    case FOLP_STEP_ADAPTIVE: take_step_adaptive(h, -1, NULL); break;
    case FOLP_STEP_MALITSKY_POCK: take_step_malitsky_pock(h); break;
    default: take_step_constant(h); break;
  }
  h->basic_time += now_sec() - t0;
}

/* ------------------------------------------------------------------------ */
/* restart scheme, sp.jl:432-927                                             */
/* ------------------------------------------------------------------------ */
static double distance_traveled(const oracle_handle* h, const double* x,
                                const double* y, const double* wp, const double* wd) {
  int64_t n = h->P.n, m = h->P.m;
  double* dxv = dalloc(n);
  double* dyv = dalloc(m);
  for (int64_t j = 0; j < n; ++j) dxv[j] = x[j] - h->last_x[j];
  for (int64_t i = 0; i < m; ++i) dyv[i] = y[i] - h->last_y[i];
  double a = weighted_norm(dxv, wp, n), b = weighted_norm(dyv, wd, m);
  free(dxv); free(dyv);
  return sqrt(a * a + b * b);
}

/* run_restart_scheme, sp.jl:688-846 */
static int run_restart_scheme(oracle_handle* h, int64_t iterations_completed,
                              const double* wp, const double* wd) {
  const folp_params* rp = &h->prm;
  int64_t n = h->P.n, m = h->P.m;
  if (!(h->count_x > 0 && h->count_y > 0)) return FOLP_RESTART_CHOICE_NO_RESTART;
  double* avg_x = dalloc(n);
  double* avg_y = dalloc(m);
  for (int64_t j = 0; j < n; ++j) avg_x[j] = h->sum_x[j] / h->w_x;
  for (int64_t i = 0; i < m; ++i) avg_y[i] = h->sum_y[i] / h->w_y;
  int64_t restart_length = h->count_x;
  int do_restart = 0;
  if ((double)restart_length >= rp->artificial_restart_threshold * (double)iterations_completed)
    do_restart = 1; /* artificial */
  int reset_to_average;
  int have_candidate = 0;
  bound_result candidate_gap = {0, 0, 0};
  double candidate_distance = 0.0;
  if (rp->restart_scheme == FOLP_NO_RESTARTS) {
    reset_to_average = 0;
  } else {
    /* compute_localized_duality_gaps, sp.jl:432-496 */
    int approx = rp->use_approximate_localized_duality_gap;
    double d_avg = distance_traveled(h, avg_x, avg_y, wp, wd);
    bound_result g_avg = bound_optimal_objective(&h->P, avg_x, avg_y, wp, wd, d_avg, 1, approx, NULL, NULL);
    double d_cur = distance_traveled(h, h->x, h->y, wp, wd);
    bound_result g_cur = bound_optimal_objective(&h->P, h->x, h->y, wp, wd, d_cur, 1, approx, NULL, NULL);
    /* should_reset_to_average, sp.jl:530-547 */
    double cur_ng = get_gap(&g_cur) / d_cur;
    double avg_ng = get_gap(&g_avg) / d_avg;
    if (rp->restart_to_current_metric == FOLP_GAP_OVER_DISTANCE_SQUARED)
      reset_to_average = (cur_ng / d_cur >= avg_ng / d_avg);
    else if (rp->restart_to_current_metric == FOLP_GAP_OVER_DISTANCE)
      reset_to_average = (cur_ng >= avg_ng);
    else
      reset_to_average = 1;
    have_candidate = 1;
    if (reset_to_average) { candidate_gap = g_avg; candidate_distance = d_avg; }
    else { candidate_gap = g_cur; candidate_distance = d_cur; }
  }
  if (!do_restart) {
    if (rp->restart_scheme == FOLP_ADAPTIVE_NORMALIZED) {
      /* should_do_adaptive_restart_normalized_duality_gap, sp.jl:549-593 */
      double pw = h->primal_weight;
      double d_last = sqrt(h->pd_last * h->pd_last * pw + h->dd_last * h->dd_last / pw);
      bound_result g_last = bound_optimal_objective(
          &h->P, h->last_x, h->last_y, wp, wd, d_last, 1,
          rp->use_approximate_localized_duality_gap, NULL, NULL);
      double ncg = get_gap(&candidate_gap) / candidate_distance;
      double nlg = get_gap(&g_last) / d_last;
      double ratio = ncg / nlg;
      if (ratio < rp->necessary_reduction_for_restart) {
        if (ratio < rp->sufficient_reduction_for_restart) do_restart = 1;
        else if (ratio > h->gap_reduction_ratio_last_trial) do_restart = 1;
      }
      h->gap_reduction_ratio_last_trial = ratio;
    } else if ((rp->restart_scheme == FOLP_ADAPTIVE_LOCALIZED ||
                rp->restart_scheme == FOLP_ADAPTIVE_DISTANCE) && !h->has_last_gap) {
      do_restart = 1;
    } else if (rp->restart_scheme == FOLP_ADAPTIVE_LOCALIZED) {
      /* should_do_localized_adaptive_restart, sp.jl:597-620 */
      double new_potential = get_gap(&candidate_gap) / (double)restart_length;
      double old_potential = h->last_gap / (double)h->last_restart_length;
      if (new_potential / old_potential < rp->necessary_reduction_for_restart) do_restart = 1;
    } else if (rp->restart_scheme == FOLP_ADAPTIVE_DISTANCE) {
      /* should_do_distance_based_adaptive_restart, sp.jl:623-648 */
      double pw = h->primal_weight;
      double d_last = sqrt(h->pd_last * h->pd_last * pw + h->dd_last * h->dd_last / pw);
      double new_potential = candidate_distance / (double)restart_length;
      double old_potential = d_last / (double)h->last_restart_length;
      if (new_potential / old_potential < rp->necessary_reduction_for_restart) do_restart = 1;
    } else if (rp->restart_scheme == FOLP_FIXED_FREQUENCY &&
               rp->restart_frequency_if_fixed <= restart_length) {
      do_restart = 1;
    }
  }
  int choice;
  if (!do_restart) {
    choice = FOLP_RESTART_CHOICE_NO_RESTART;
  } else {
    if (reset_to_average) {
      memcpy(h->x, avg_x, sizeof(double) * (size_t)n);
      memcpy(h->y, avg_y, sizeof(double) * (size_t)m);
    }
    /* reset_solution_weighted_average, sp.jl:238-250 */
    memset(h->sum_x, 0, sizeof(double) * (size_t)n);
    memset(h->sum_y, 0, sizeof(double) * (size_t)m);
    h->count_x = h->count_y = 0;
    h->w_x = h->w_y = 0.0;
    /* update_last_restart_info, sp.jl:893-927 */
    double* dxv = dalloc(n);
    double* dyv = dalloc(m);
    for (int64_t j = 0; j < n; ++j) dxv[j] = avg_x[j] - h->last_x[j];
    for (int64_t i = 0; i < m; ++i) dyv[i] = avg_y[i] - h->last_y[i];
    h->pd_last = weighted_norm(dxv, wp, n) / sqrt(h->primal_weight);
    h->dd_last = weighted_norm(dyv, wd, m) * sqrt(h->primal_weight);
    free(dxv); free(dyv);
    memcpy(h->last_x, h->x, sizeof(double) * (size_t)n);
    memcpy(h->last_y, h->y, sizeof(double) * (size_t)m);
    h->last_restart_length = restart_length;
    h->has_last_gap = have_candidate;
    h->last_gap = have_candidate ? get_gap(&candidate_gap) : 0.0;
    choice = reset_to_average ? FOLP_RESTART_CHOICE_RESTART_TO_AVERAGE
                              : FOLP_RESTART_CHOICE_WEIGHTED_AVERAGE_RESET;
  }
  free(avg_x); free(avg_y);
  return choice;
}

/* compute_new_primal_weight, sp.jl:862-891 */
static double new_primal_weight(const oracle_handle* h, double pw, double smoothing) {
  const double eps = 2.220446049250313e-16;
  if (h->pd_last > eps && h->dd_last > eps) {
    double est = h->dd_last / h->pd_last;
    double lpw = smoothing * log(est) + (1 - smoothing) * log(pw);
    return exp(lpw);
  }
  return pw;
}

/* ------------------------------------------------------------------------ */
/* the loop, pdhg.jl:782-1049                                                */
/* ------------------------------------------------------------------------ */
int oracle_create(const folp_problem* fp, const folp_params* prm, oracle_handle** out) {
  if (!fp || !prm || !out) return FOLP_INVALID_ARGUMENT;
  oracle_handle* h = (oracle_handle*)calloc(1, sizeof(*h));
  qp_from_folp(fp, 0, &h->P);
  qp_from_folp(fp, 1, &h->O);
  int64_t n = h->P.n, m = h->P.m;
  h->D = dalloc(n); h->E = dalloc(m);
  for (int64_t j = 0; j < n; ++j) h->D[j] = fp->variable_rescaling ? fp->variable_rescaling[j] : 1.0;
  for (int64_t i = 0; i < m; ++i) h->E[i] = fp->constraint_rescaling ? fp->constraint_rescaling[i] : 1.0;
  h->cache[0] = fp->l_inf_norm_primal_linear_objective;
  h->cache[1] = fp->l_inf_norm_primal_right_hand_side;
  h->cache[2] = fp->l2_norm_primal_linear_objective;
  h->cache[3] = fp->l2_norm_primal_right_hand_side;
  h->prm = *prm;
  h->p_is_lp = qp_is_lp(&h->P);
  if (prm->step_size_policy == FOLP_STEP_MALITSKY_POCK && !h->p_is_lp) {
    /* pdhg.jl:560-565 */
    oracle_destroy(h);
    return FOLP_UNSUPPORTED;
  }
  /* :805-819 */
  h->x = dzeros(n); h->y = dzeros(m); h->dx = dzeros(n); h->dy = dzeros(m);
  h->aty = dzeros(n);
  h->sum_x = dzeros(n); h->sum_y = dzeros(m);
  h->step_size = prm->initial_step_size;     /* :821-839 computed by the host */
  h->primal_weight = prm->initial_primal_weight; /* :847-857 */
  h->kkt_passes = prm->initial_kkt_passes;
  h->ratio_step_sizes = 1.0;
  /* create_last_restart_info, sp.jl:199-213 */
  h->last_x = dzeros(n); h->last_y = dzeros(m);
  h->has_last_gap = 0; h->last_restart_length = 1;
  h->pd_last = 0.0; h->dd_last = 0.0; h->gap_reduction_ratio_last_trial = 1.0;
  h->iteration = 0; h->need_step = 0; h->terminated = 0;
  h->start_time = now_sec(); h->basic_time = 0.0;
  *out = h;
  return FOLP_OK;
}

void oracle_destroy(oracle_handle* h) {
  if (!h) return;
  qp_free(&h->P); qp_free(&h->O);
  free(h->D); free(h->E);
  free(h->x); free(h->y); free(h->dx); free(h->dy); free(h->aty);
  free(h->sum_x); free(h->sum_y); free(h->last_x); free(h->last_y);
  free(h);
}

double oracle_basic_seconds(const oracle_handle* h) { return h->basic_time; }

/* avg used by the evaluation block, pdhg.jl:902-910. Returns 1 if freshly
 * allocated. */
static int current_average(const oracle_handle* h, double** ax, double** ay) {
  if (h->numerical_error || h->count_x == 0 || h->count_y == 0) {
    *ax = h->x; *ay = h->y;
    return 0;
  }
  int64_t n = h->P.n, m = h->P.m;
  *ax = dalloc(n); *ay = dalloc(m);
  for (int64_t j = 0; j < n; ++j) (*ax)[j] = h->sum_x[j] / h->w_x;
  for (int64_t i = 0; i < m; ++i) (*ay)[i] = h->sum_y[i] / h->w_y;
  return 1;
}

int oracle_run(oracle_handle* h, folp_eval* out) {
  if (!h || !out) return FOLP_INVALID_ARGUMENT;
  if (h->terminated) { *out = h->last_eval; return FOLP_OK; }
  const folp_params* prm = &h->prm;
  int64_t n = h->P.n, m = h->P.m;
  for (;;) {
    if (h->need_step) { take_step(h); h->need_step = 0; } /* :1044 */
    h->iteration += 1;                                    /* :887 */
    int64_t iteration = h->iteration;
    if (!((iteration - 1) % prm->termination_evaluation_frequency == 0 ||
          iteration == (int64_t)prm->iteration_limit + 1 || iteration <= 10 ||
          h->numerical_error)) {                          /* :892-895 */
      h->need_step = 1;
      continue;
    }
    h->kkt_passes += 2.0;                                 /* :899 */
    double *avg_x, *avg_y;
    int owned = current_average(h, &avg_x, &avg_y);
    /* evaluate_unscaled_iteration_stats, isu.jl:413-451 */
    double* ox = dalloc(n);
    double* oy = dalloc(m);
    for (int64_t j = 0; j < n; ++j) ox[j] = avg_x[j] / h->D[j];
    for (int64_t i = 0; i < m; ++i) oy[i] = avg_y[i] / h->E[i];
    folp_eval e;
    memset(&e, 0, sizeof(e));
    e.iteration_number = (int32_t)(iteration - 1);
    e.cumulative_kkt_matrix_passes = h->kkt_passes;
    e.cumulative_time_sec = now_sec() - h->start_time;
    convergence_information(&h->O, h->cache, ox, oy,
                            prm->eps_optimal_absolute / prm->eps_optimal_relative,
                            FOLP_POINT_TYPE_AVERAGE_ITERATE, &e);
    infeasibility_information(&h->O, ox, oy, &e);
    free(ox); free(oy);
    e.step_size = h->step_size;
    e.primal_weight = h->primal_weight;
    e.time_spent_doing_basic_algorithm = h->basic_time;   /* :929 */
    /* define_norms, :265-276 */
    double* wp = dalloc(n);
    double* wd = dalloc(m);
    double wpv = 1 / h->step_size * h->primal_weight;
    double wdv = 1 / h->step_size / h->primal_weight;
    for (int64_t j = 0; j < n; ++j) wp[j] = wpv;
    for (int64_t i = 0; i < m; ++i) wd[i] = wdv;
    { /* update_objective_bound_estimates, sp.jl:1015-1047 */
      double rp_ = jl_max(1e-8, weighted_norm(avg_x, wp, n));
      double rd_ = jl_max(1e-8, weighted_norm(avg_y, wd, m));
      double* wp2 = dalloc(n);
      double* wd2 = dalloc(m);
      for (int64_t j = 0; j < n; ++j) wp2[j] = wp[j] / (rp_ * rp_);
      for (int64_t i = 0; i < m; ++i) wd2[i] = wd[i] / (rd_ * rd_);
      bound_result br = bound_optimal_objective(&h->P, avg_x, avg_y, wp2, wd2, 1.0, 0, 0, NULL, NULL);
      e.lagrangian_value = br.lagrangian_value;
      e.estimated_lower_bound = br.lower_bound_value;
      e.estimated_upper_bound = br.upper_bound_value;
      free(wp2); free(wd2);
    }
    int reason = check_termination(prm, h->cache, &e);    /* :947 */
    if (h->numerical_error && reason == 0) reason = FOLP_TERMINATION_REASON_NUMERICAL_ERROR;
    e.termination_reason = reason;
    e.numerical_error = h->numerical_error;
    e.total_number_iterations = h->total_number_iterations;
    if (reason != 0) {                                    /* :972-993 */
      if (owned) { free(avg_x); free(avg_y); }
      free(wp); free(wd);
      h->terminated = 1;
      e.restart_used = FOLP_RESTART_CHOICE_UNSPECIFIED;
      h->last_eval = e; *out = e;
      return FOLP_OK;
    }
    if (owned) { free(avg_x); free(avg_y); }
    e.restart_used = run_restart_scheme(h, iteration - 1, wp, wd); /* :995 */
    free(wp); free(wd);
    if (e.restart_used != FOLP_RESTART_CHOICE_NO_RESTART) {        /* :1009-1017 */
      h->primal_weight = new_primal_weight(h, h->primal_weight, prm->primal_weight_update_smoothing);
      h->ratio_step_sizes = 1.0;
    }
    if (e.restart_used == FOLP_RESTART_CHOICE_RESTART_TO_AVERAGE)  /* :1018-1022 */
      csc_tmul(&h->P.A, h->y, h->aty);
    h->need_step = 1;
    h->last_eval = e; *out = e;
    return FOLP_OK;
  }
}

int oracle_get_solution(oracle_handle* h, int which, int unscaled, double* x_out, double* y_out) {
  if (!h) return FOLP_INVALID_ARGUMENT;
  int64_t n = h->P.n, m = h->P.m;
  double *ax, *ay;
  int owned = 0;
  if (which == 0) owned = current_average(h, &ax, &ay);
  else { ax = h->x; ay = h->y; }
  if (x_out) for (int64_t j = 0; j < n; ++j) x_out[j] = unscaled ? ax[j] / h->D[j] : ax[j];
  if (y_out) for (int64_t i = 0; i < m; ++i) y_out[i] = unscaled ? ay[i] / h->E[i] : ay[i];
  if (owned) { free(ax); free(ay); }
  return FOLP_OK;
}

int oracle_solve(oracle_handle* h, folp_eval* evals, int64_t max_evals, int64_t* num_evals,
                 int32_t* termination_reason, int32_t* iteration_count, double* x_out,
                 double* y_out) {
  if (!h) return FOLP_INVALID_ARGUMENT;
  int64_t cnt = 0;
  folp_eval e;
  for (;;) {
    int rc = oracle_run(h, &e);
    if (rc) return rc;
    /* pdhg.jl:958-960 */
    if (evals && (h->prm.record_iteration_stats || e.termination_reason != 0)) {
      if (cnt < max_evals) evals[cnt] = e;
      else if (max_evals > 0) evals[max_evals - 1] = e;
      if (cnt < max_evals) cnt++;
    }
    if (e.termination_reason != 0) break;
  }
  if (num_evals) *num_evals = cnt;
  if (termination_reason) *termination_reason = e.termination_reason;
  if (iteration_count) *iteration_count = e.iteration_number;
  return oracle_get_solution(h, 0, 1, x_out, y_out);
}

/* One inner attempt of take_step(::AdaptiveStepsizeParams) (the body of the
 * while loop, pdhg.jl:662-729). h->step_size carries the trial step between
 * attempts; the averaging weight stays the step size at entry (:512). */
static void single_adaptive_attempt(oracle_handle* h) {
  if (!h->in_retry) h->entry_weight = h->step_size;
  double step_size = h->step_size;
  h->total_number_iterations += 1;
  double* np = next_primal_solution(h, step_size);
  double *nd, *naty;
  next_dual_solution(h, np, step_size, 1.0, &nd, &naty);
  double interaction, movement;
  interaction_and_movement(h, np, nd, naty, &interaction, &movement);
  h->last_interaction = interaction; h->last_movement = movement;
  h->kkt_passes += 1;
  if (movement == 0.0) {
    h->numerical_error = 1;
    free(np); free(nd); free(naty);
    return;
  }
  double limit = interaction > 0 ? movement / interaction : INFINITY;
  if (step_size <= limit) {
    update_solution(h, np, nd, naty, h->entry_weight);
    h->in_retry = 0;
  } else {
    free(np); free(nd); free(naty);
    h->in_retry = 1;
  }
  double k1 = (double)(h->total_number_iterations + 1);
  double first_term = (1 - pow(k1, -h->prm.reduction_exponent)) * limit;
  double second_term = (1 + pow(k1, -h->prm.growth_exponent)) * step_size;
  h->step_size = jl_min(first_term, second_term);
}

int oracle_debug_attempts(oracle_handle* h, int64_t attempts) {
  if (!h) return FOLP_INVALID_ARGUMENT;
  for (int64_t k = 0; k < attempts && !h->numerical_error; ++k) {
    if (h->prm.step_size_policy == FOLP_STEP_ADAPTIVE) single_adaptive_attempt(h);
    else take_step(h);
  }
  return FOLP_OK;
}

int oracle_debug_state(oracle_handle* h, double* x, double* y, double* dual_product,
                       double* sum_x, double* sum_y, folp_debug_scalars* s) {
  if (!h) return FOLP_INVALID_ARGUMENT;
  size_t nb = sizeof(double) * (size_t)h->P.n, mb = sizeof(double) * (size_t)h->P.m;
  if (x) memcpy(x, h->x, nb);
  if (y) memcpy(y, h->y, mb);
  if (dual_product) memcpy(dual_product, h->aty, nb);
  if (sum_x) memcpy(sum_x, h->sum_x, nb);
  if (sum_y) memcpy(sum_y, h->sum_y, mb);
  if (s) {
    memset(s, 0, sizeof(*s));
    s->step_size = h->step_size; s->primal_weight = h->primal_weight;
    s->cumulative_kkt_passes = h->kkt_passes;
    s->sum_primal_solution_weights = h->w_x; s->sum_dual_solution_weights = h->w_y;
    s->total_number_iterations = h->total_number_iterations;
    s->iterations_completed = h->count_x; /* since last restart */
    s->sum_primal_solutions_count = h->count_x; s->sum_dual_solutions_count = h->count_y;
    s->numerical_error = h->numerical_error;
    s->last_interaction = h->last_interaction; s->last_movement = h->last_movement;
  }
  return FOLP_OK;
}

int oracle_debug_set_state(oracle_handle* h, const double* x, const double* y,
                           double step_size, double primal_weight) {
  if (!h) return FOLP_INVALID_ARGUMENT;
  if (x) memcpy(h->x, x, sizeof(double) * (size_t)h->P.n);
  if (y) { memcpy(h->y, y, sizeof(double) * (size_t)h->P.m); csc_tmul(&h->P.A, h->y, h->aty); }
  if (step_size > 0) h->step_size = step_size;
  if (primal_weight > 0) h->primal_weight = primal_weight;
  return FOLP_OK;
}

int oracle_debug_spmv(oracle_handle* h, int transpose, const double* in, double* out) {
  if (!h) return FOLP_INVALID_ARGUMENT;
  if (transpose) csc_tmul(&h->P.A, in, out);
  else csc_mul(&h->P.A, in, out);
  return FOLP_OK;
}

/* ------------------------------------------------------------------------ */
/* unit-level entry points                                                   */
/* ------------------------------------------------------------------------ */
int oracle_trust_region(int64_t len, const double* center_point, const double* objective_vector,
                        const double* lb, const double* ub, const double* norm_weights,
                        double target_radius, int solve_approximately, double* solution_out,
                        double* value_out) {
  *value_out = tr_solve(len, center_point, objective_vector, lb, ub, norm_weights,
                        target_radius, solve_approximately, solution_out);
  return FOLP_OK;
}

int oracle_bound_optimal_objective(const folp_problem* fp, const double* x, const double* y,
                                   const double* wp, const double* wd, double radius,
                                   int norm_kind, int approx, double* out3, double* x_tr,
                                   double* y_tr) {
  qp_t p;
  qp_from_folp(fp, 0, &p);
  bound_result r = bound_optimal_objective(&p, x, y, wp, wd, radius, norm_kind, approx, x_tr, y_tr);
  out3[0] = r.lagrangian_value; out3[1] = r.lower_bound_value; out3[2] = r.upper_bound_value;
  qp_free(&p);
  return FOLP_OK;
}

static void cache_from_qp(const qp_t* p, double cache[4]) {
  /* cached_quadratic_program_info, term.jl:151-158 */
  cache[0] = norminf(p->c, p->n); cache[1] = norminf(p->b, p->m);
  cache[2] = norm2(p->c, p->n); cache[3] = norm2(p->b, p->m);
}

int oracle_iteration_stats(const folp_problem* fp, const double* x, const double* y,
                           const double* xray, const double* yray, double eps_abs,
                           double eps_rel, folp_eval* out) {
  qp_t p;
  qp_from_folp(fp, 0, &p);
  double cache[4];
  cache_from_qp(&p, cache);
  convergence_information(&p, cache, x, y, eps_abs / eps_rel, out->candidate_type, out);
  infeasibility_information(&p, xray, yray, out);
  qp_free(&p);
  return FOLP_OK;
}

int oracle_dual_stats(const folp_problem* fp, const double* x, const double* y,
                      double* dual_objective, double* dres, double* rc) {
  qp_t p;
  qp_from_folp(fp, 0, &p);
  *dual_objective = dual_stats(&p, x, y, dres, rc);
  qp_free(&p);
  return FOLP_OK;
}
double oracle_max_primal_violation(const folp_problem* fp, const double* x) {
  qp_t p;
  qp_from_folp(fp, 0, &p);
  double* r = dalloc(p.m + 2 * p.n);
  primal_residual(&p, x, r);
  double v = norminf(r, p.m + 2 * p.n);
  free(r); qp_free(&p);
  return v;
}
double oracle_primal_obj(const folp_problem* fp, const double* x) {
  qp_t p;
  qp_from_folp(fp, 0, &p);
  double v = primal_obj(&p, x);
  qp_free(&p);
  return v;
}
double oracle_lagrangian_value(const folp_problem* fp, const double* x, const double* y) {
  qp_t p;
  qp_from_folp(fp, 0, &p);
  double v = lagrangian_value(&p, x, y);
  qp_free(&p);
  return v;
}
/* select_initial_primal_weight, sp.jl:1049-1075 */
double oracle_select_initial_primal_weight(const folp_problem* fp, const double* wp,
                                           const double* wd, double primal_importance) {
  double rhs = weighted_norm(fp->right_hand_side, wd, fp->num_constraints);
  double obj = weighted_norm(fp->objective_vector, wp, fp->num_variables);
  if (obj > 0.0 && rhs > 0.0) return primal_importance * (obj / rhs);
  return primal_importance;
}
int oracle_check_termination(const folp_params* params, const folp_problem* fp,
                             const folp_eval* stats) {
  double cache[4] = {fp->l_inf_norm_primal_linear_objective, fp->l_inf_norm_primal_right_hand_side,
                     fp->l2_norm_primal_linear_objective, fp->l2_norm_primal_right_hand_side};
  return check_termination(params, cache, stats);
}

/* ------------------------------------------------------------------------ */
/* rescaling, pre.jl:99-113, 358-573, 631-687                                */
/* ------------------------------------------------------------------------ */
/* maximum(abs, matrix, dims): dims = 1 -> per column (n), 2 -> per row (m) */
static void max_abs(const csc_t* A, int dimension, double* out) {
  int64_t len = dimension == 1 ? A->n : A->m;
  for (int64_t i = 0; i < len; ++i) out[i] = 0.0;
  for (int64_t j = 0; j < A->n; ++j)
    for (int64_t k = A->colptr[j]; k < A->colptr[j + 1]; ++k) {
      int64_t idx = dimension == 1 ? j : A->rowval[k];
      double a = fabs(A->nzval[k]);
      if (a > out[idx]) out[idx] = a;
    }
}
/* l2_norm, pre.jl:99-113 */
static void l2_norm(const csc_t* A, int dimension, double* out) {
  int64_t len = dimension == 1 ? A->n : A->m;
  double* sf = dalloc(len);
  max_abs(A, dimension, sf);
  for (int64_t i = 0; i < len; ++i) if (sf[i] == 0.0) sf[i] = 1.0;
  for (int64_t i = 0; i < len; ++i) out[i] = 0.0;
  for (int64_t j = 0; j < A->n; ++j)
    for (int64_t k = A->colptr[j]; k < A->colptr[j + 1]; ++k) {
      int64_t idx = dimension == 1 ? j : A->rowval[k];
      double t = dimension == 1 ? A->nzval[k] * (1 / sf[idx]) : (1 / sf[idx]) * A->nzval[k];
      out[idx] += t * t;
    }
  for (int64_t i = 0; i < len; ++i) out[i] = sf[i] * sqrt(out[i]);
  free(sf);
}
typedef struct {
  int64_t m, n;
  csc_t A, Q;
  double *c, *l, *u, *b;
} scale_view;
/* scale_problem, pre.jl:555-573 */
static void scale_problem(scale_view* p, const double* con, const double* var) {
  for (int64_t j = 0; j < p->n; ++j) p->c[j] /= var[j];
  for (int64_t j = 0; j < p->n; ++j)
    for (int64_t k = p->Q.colptr[j]; k < p->Q.colptr[j + 1]; ++k)
      p->Q.nzval[k] = ((1 / var[p->Q.rowval[k]]) * p->Q.nzval[k]) * (1 / var[j]);
  for (int64_t j = 0; j < p->n; ++j) p->u[j] *= var[j];
  for (int64_t j = 0; j < p->n; ++j) p->l[j] *= var[j];
  for (int64_t i = 0; i < p->m; ++i) p->b[i] /= con[i];
  for (int64_t j = 0; j < p->n; ++j)
    for (int64_t k = p->A.colptr[j]; k < p->A.colptr[j + 1]; ++k)
      p->A.nzval[k] = ((1 / con[p->A.rowval[k]]) * p->A.nzval[k]) * (1 / var[j]);
}
static int view_is_lp(const scale_view* p) {
  for (int64_t k = 0; k < p->Q.nnz; ++k) if (p->Q.nzval[k] != 0.0) return 0;
  return 1;
}
/* ruiz_rescaling, pre.jl:412-477 */
static void ruiz_rescaling(scale_view* p, int iterations, int pnorm, double* cum_con, double* cum_var) {
  int64_t m = p->m, n = p->n;
  for (int64_t i = 0; i < m; ++i) cum_con[i] = 1.0;
  for (int64_t j = 0; j < n; ++j) cum_var[j] = 1.0;
  double* var = dalloc(n); double* con = dalloc(m);
  double* t1 = dalloc(n); double* t2 = dalloc(n);
  for (int it = 0; it < iterations; ++it) {
    if (pnorm == 0) {
      max_abs(&p->A, 1, t1); max_abs(&p->Q, 1, t2);
      for (int64_t j = 0; j < n; ++j) var[j] = sqrt(jl_max(t1[j], t2[j]));
    } else {
      l2_norm(&p->A, 1, t1); l2_norm(&p->Q, 1, t2);
      for (int64_t j = 0; j < n; ++j) var[j] = sqrt(sqrt(t1[j] * t1[j] + t2[j] * t2[j]));
    }
    for (int64_t j = 0; j < n; ++j) if (var[j] == 0.0) var[j] = 1.0;
    if (m > 0) {
      if (pnorm == 0) {
        max_abs(&p->A, 2, con);
        for (int64_t i = 0; i < m; ++i) con[i] = sqrt(con[i]);
      } else {
        l2_norm(&p->A, 2, con);
        double target = view_is_lp(p) ? sqrt((double)n / (double)m)
                                      : sqrt((double)n / (double)(m + n));
        for (int64_t i = 0; i < m; ++i) con[i] = sqrt(con[i] / target);
      }
      for (int64_t i = 0; i < m; ++i) if (con[i] == 0.0) con[i] = 1.0;
    }
    scale_problem(p, con, var);
    for (int64_t i = 0; i < m; ++i) cum_con[i] *= con[i];
    for (int64_t j = 0; j < n; ++j) cum_var[j] *= var[j];
  }
  free(var); free(con); free(t1); free(t2);
}
/* l2_norm_rescaling, pre.jl:358-372 */
static void l2_norm_rescaling(scale_view* p, double* con, double* var) {
  l2_norm(&p->A, 2, con); l2_norm(&p->A, 1, var);
  for (int64_t i = 0; i < p->m; ++i) if (con[i] == 0.0) con[i] = 1.0;
  for (int64_t j = 0; j < p->n; ++j) if (var[j] == 0.0) var[j] = 1.0;
  for (int64_t j = 0; j < p->n; ++j) var[j] = sqrt(var[j]);
  for (int64_t i = 0; i < p->m; ++i) con[i] = sqrt(con[i]);
  scale_problem(p, con, var);
}
/* pock_chambolle_rescaling, pre.jl:508-539. mapreduce over a sparse matrix
 * with dims also folds f(0) for every structural zero: |0|^0 == 1. */
static void pock_chambolle_rescaling(scale_view* p, double alpha, double* con, double* var) {
  int64_t m = p->m, n = p->n;
  int64_t* row_count = (int64_t*)calloc((size_t)(m > 0 ? m : 1), sizeof(int64_t));
  for (int64_t j = 0; j < n; ++j) var[j] = 0.0;
  for (int64_t i = 0; i < m; ++i) con[i] = 0.0;
  double zero_col = pow(0.0, 2 - alpha), zero_row = pow(0.0, alpha);
  for (int64_t j = 0; j < n; ++j) {
    for (int64_t k = p->A.colptr[j]; k < p->A.colptr[j + 1]; ++k) {
      double a = fabs(p->A.nzval[k]);
      var[j] += pow(a, 2 - alpha);
      con[p->A.rowval[k]] += pow(a, alpha);
      row_count[p->A.rowval[k]] += 1;
    }
    var[j] += zero_col * (double)(m - (p->A.colptr[j + 1] - p->A.colptr[j]));
  }
  for (int64_t i = 0; i < m; ++i) con[i] += zero_row * (double)(n - row_count[i]);
  for (int64_t j = 0; j < n; ++j) { var[j] = sqrt(var[j]); if (var[j] == 0.0) var[j] = 1.0; }
  for (int64_t i = 0; i < m; ++i) { con[i] = sqrt(con[i]); if (con[i] == 0.0) con[i] = 1.0; }
  free(row_count);
  scale_problem(p, con, var);
}

int oracle_l2_norm(int64_t m, int64_t n, const int64_t* colptr, const int64_t* rowval,
                   const double* nzval, int dimension, double* out) {
  csc_t A;
  memset(&A, 0, sizeof(A));
  A.m = m; A.n = n; A.nnz = colptr[n];
  A.colptr = (int64_t*)colptr; A.rowval = (int64_t*)rowval; A.nzval = (double*)nzval;
  l2_norm(&A, dimension, out);
  return FOLP_OK;
}

/* rescale_problem, pre.jl:631-687 */
int oracle_rescale_problem(int64_t m, int64_t n, const int64_t* colptr, const int64_t* rowval,
                           double* nzval, const int64_t* q_colptr, const int64_t* q_rowval,
                           double* q_nzval, double* c, double* l, double* u, double* b,
                           int l_inf_ruiz_iterations, int ruiz_p, int l2_rescaling,
                           double pock_chambolle_alpha, double* constraint_rescaling,
                           double* variable_rescaling) {
  scale_view p;
  memset(&p, 0, sizeof(p));
  p.m = m; p.n = n;
  p.A.m = m; p.A.n = n; p.A.nnz = colptr[n];
  p.A.colptr = (int64_t*)colptr; p.A.rowval = (int64_t*)rowval; p.A.nzval = nzval;
  int64_t* zero_colptr = NULL;
  p.Q.m = n; p.Q.n = n;
  if (q_colptr) {
    p.Q.nnz = q_colptr[n]; p.Q.colptr = (int64_t*)q_colptr;
    p.Q.rowval = (int64_t*)q_rowval; p.Q.nzval = q_nzval;
  } else {
    zero_colptr = (int64_t*)calloc((size_t)n + 1, sizeof(int64_t));
    p.Q.nnz = 0; p.Q.colptr = zero_colptr;
  }
  p.c = c; p.l = l; p.u = u; p.b = b;
  for (int64_t i = 0; i < m; ++i) constraint_rescaling[i] = 1.0;
  for (int64_t j = 0; j < n; ++j) variable_rescaling[j] = 1.0;
  double* con = dalloc(m); double* var = dalloc(n);
  if (l_inf_ruiz_iterations > 0) {
    ruiz_rescaling(&p, l_inf_ruiz_iterations, ruiz_p, con, var);
    for (int64_t i = 0; i < m; ++i) constraint_rescaling[i] *= con[i];
    for (int64_t j = 0; j < n; ++j) variable_rescaling[j] *= var[j];
  }
  if (l2_rescaling) {
    l2_norm_rescaling(&p, con, var);
    for (int64_t i = 0; i < m; ++i) constraint_rescaling[i] *= con[i];
    for (int64_t j = 0; j < n; ++j) variable_rescaling[j] *= var[j];
  }
  if (pock_chambolle_alpha >= 0.0) {
    pock_chambolle_rescaling(&p, pock_chambolle_alpha, con, var);
    for (int64_t i = 0; i < m; ++i) constraint_rescaling[i] *= con[i];
    for (int64_t j = 0; j < n; ++j) variable_rescaling[j] *= var[j];
  }
  free(con); free(var); free(zero_colptr);
  return FOLP_OK;
}
