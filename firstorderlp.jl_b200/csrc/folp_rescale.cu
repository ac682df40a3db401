// folp_rescale.cu -- rescale_problem (src/preprocess.jl:631-687) on the device: Ruiz
// (preprocess.jl:412-477), l2-norm (:358-372) and Pock-Chambolle (:508-539) rescaling of a
// QuadraticProgrammingProblem held in Julia's CSC layout. SURVEY.md section 8f-1: the step
// immediately before the PDHG loop, k+2 full passes over the matrix that the reference does
// serially on the host, building a new matrix each time.
//
// Layout in HBM: the caller's CSC of A (int32 indices after an on-device narrowing), the
// column of every entry (colof), and the CSR view of the same values as a permutation
// (rowptr, perm: the k-th entry of row i in ascending column order sits at CSC position
// perm[k]) -- one copy of the values, scaled in place, read through either view.
// Arithmetic: element-wise operations are the reference's own, in its order
// (((1/con_i) * a_ij) * (1/var_j), c_j / var_j, ...; -fmad=false); a row's or a column's sum
// runs sequentially in one thread in the order the reference's loops visit it (ascending row
// inside a column, ascending column inside a row), so the result is bit-identical to the CPU
// restatement for every row and column of up to kLongSegment entries; longer ones (PageRank's
// dense row) are summed by a block-wide fixed-shape tree instead.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <thread>
#include <vector>

#include "folp_internal.cuh"

namespace folp {
void set_create_error(const std::string& msg);  // folp_api.cu: what folp_last_error(NULL) returns
// folp_api.cu: stable CSC -> CSR transposition returning the permutation (no CUDA call)
bool csr_permutation(int64_t n, int64_t m, int64_t nnz, const int64_t* colptr, const int64_t* rowval,
                     int base, std::vector<int>* rowptr, std::vector<int>* perm);
}  // namespace folp

using namespace folp;

namespace {

constexpr int kThreads = 256;
constexpr int kLongSegment = 2048;

enum Stat : int { kMaxAbs = 0, kSumSqScaled = 1, kSumPow = 2 };

__device__ __forceinline__ double stat_term(int stat, double v, double inv_sf, double e) {
  const double a = fabs(v);
  if (stat == kMaxAbs) return a;
  if (stat == kSumSqScaled) {
    const double t = v * inv_sf;  // preprocess.jl:107-110
    return t * t;
  }
  // |a|^e, preprocess.jl:520-533. Exponents 1, 2 and 0 are exact in every libm; anything else goes
  // through the device pow (<= 2 ulp from the host's).
  if (e == 1.0) return a;
  if (e == 2.0) return a * a;
  if (e == 0.0) return 1.0;
  return pow(a, e);
}

// One thread per segment (a column through ptr/vals, or a row through ptr/perm/vals), entries
// visited in storage order. Segments longer than kLongSegment are left to k_seg_stat_long.
template <bool PERM>
__global__ void __launch_bounds__(kThreads) k_seg_stat(const int* __restrict__ ptr,
                                                       const int* __restrict__ perm,
                                                       const double* __restrict__ vals, int nseg,
                                                       int stat, const double* __restrict__ sf,
                                                       double e, double* __restrict__ out) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nseg) return;
  const int k0 = ptr[s], k1 = ptr[s + 1];
  if (k1 - k0 > kLongSegment) return;
  const double inv_sf = stat == kSumSqScaled ? 1 / sf[s] : 1.0;
  double acc = 0.0;
  for (int k = k0; k < k1; ++k) {
    const double v = vals[PERM ? perm[k] : k];
    const double t = stat_term(stat, v, inv_sf, e);
    if (stat == kMaxAbs) acc = t > acc ? t : acc;
    else acc += t;
  }
  out[s] = acc;
}

// One block per long segment: threads stride the entries, fixed-shape tree.
template <bool PERM>
__global__ void __launch_bounds__(kThreads) k_seg_stat_long(const int* __restrict__ ptr,
                                                            const int* __restrict__ perm,
                                                            const double* __restrict__ vals,
                                                            const int* __restrict__ long_ids, int stat,
                                                            const double* __restrict__ sf, double e,
                                                            double* __restrict__ out) {
  __shared__ double sh[32];
  const int s = long_ids[blockIdx.x];
  const int k0 = ptr[s], k1 = ptr[s + 1];
  const double inv_sf = stat == kSumSqScaled ? 1 / sf[s] : 1.0;
  double acc = 0.0;
  for (int k = k0 + threadIdx.x; k < k1; k += blockDim.x) {
    const double v = vals[PERM ? perm[k] : k];
    const double t = stat_term(stat, v, inv_sf, e);
    if (stat == kMaxAbs) acc = t > acc ? t : acc;
    else acc += t;
  }
  const double r = stat == kMaxAbs ? block_reduce<true>(acc, sh) : block_reduce<false>(acc, sh);
  if (threadIdx.x == 0) out[s] = r;
}

// colof[k] = column of CSC entry k (binary search in colptr)
__global__ void __launch_bounds__(kThreads) k_col_of(const int* __restrict__ colptr, int ncols, int nnz,
                                                     int* __restrict__ colof) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t k = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; k < nnz; k += stride) {
    int lo = 0, hi = ncols;  // largest j with colptr[j] <= k
    while (hi - lo > 1) {
      const int mid = lo + (hi - lo) / 2;
      if (colptr[mid] <= k) lo = mid;
      else hi = mid;
    }
    colof[k] = lo;
  }
}

__global__ void __launch_bounds__(kThreads) k_narrow(const int64_t* __restrict__ in, int base, int64_t len,
                                                     int* __restrict__ out) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t k = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; k < len; k += stride)
    out[k] = static_cast<int>(in[k] - base);
}

// scale_problem, preprocess.jl:555-573: a_ij <- ((1/r_i) * a_ij) * (1/v_j). For A: r = constraint
// rescaling, v = variable rescaling; for Q both are the variable rescaling.
__global__ void __launch_bounds__(kThreads) k_scale_entries(const int* __restrict__ rowval,
                                                            const int* __restrict__ colof, int64_t nnz,
                                                            const double* __restrict__ r,
                                                            const double* __restrict__ v,
                                                            double* __restrict__ vals) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t k = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; k < nnz; k += stride)
    vals[k] = ((1 / r[rowval[k]]) * vals[k]) * (1 / v[colof[k]]);
}

enum VecOp : int {
  kDivBy = 0,        // a /= b
  kMulBy,            // a *= b
  kSqrtMax,          // a = sqrt(max(a, b)); 0 -> 1      (Ruiz, p = Inf: variables)
  kSqrt,             // a = sqrt(a); 0 -> 1               (Ruiz, p = Inf: constraints)
  kSqrtSqrtSumSq,    // a = sqrt(sqrt(a*a + b*b)); 0 -> 1 (Ruiz, p = 2: variables)
  kSqrtDiv,          // a = sqrt(a / s); 0 -> 1           (Ruiz, p = 2: constraints)
  kL2Finish,         // a = b * sqrt(a)                   (l2_norm: sf * sqrt(sum))
  kZeroToOne,        // a == 0 -> 1
  kZeroToOneSqrt,    // a == 0 -> 1, then sqrt            (l2_norm_rescaling)
  kAddCountSqrt,     // a = sqrt(a + s * (t - len)); 0 -> 1 (Pock-Chambolle; len from ptr)
  kFill              // a = s
};
__global__ void __launch_bounds__(kThreads) k_vec(int op, double* __restrict__ a,
                                                  const double* __restrict__ b, int64_t len, double s,
                                                  double t, const int* __restrict__ ptr) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < len; i += stride) {
    double x = a[i];
    switch (op) {
      case kDivBy: x = x / b[i]; break;
      case kMulBy: x = x * b[i]; break;
      case kSqrtMax: {
        const double y = b[i];
        x = sqrt(x != x ? x : (y != y ? y : (x > y ? x : y)));
        if (x == 0.0) x = 1.0;
        break;
      }
      case kSqrt: x = sqrt(x); if (x == 0.0) x = 1.0; break;
      case kSqrtSqrtSumSq: x = sqrt(sqrt(x * x + b[i] * b[i])); if (x == 0.0) x = 1.0; break;
      case kSqrtDiv: x = sqrt(x / s); if (x == 0.0) x = 1.0; break;
      case kL2Finish: x = b[i] * sqrt(x); break;
      case kZeroToOne: if (x == 0.0) x = 1.0; break;
      case kZeroToOneSqrt: if (x == 0.0) x = 1.0; x = sqrt(x); break;
      case kAddCountSqrt:
        x = x + s * static_cast<double>(static_cast<int64_t>(t) - (ptr[i + 1] - ptr[i]));
        x = sqrt(x);
        if (x == 0.0) x = 1.0;
        break;
      case kFill: x = s; break;
    }
    a[i] = x;
  }
}

struct Rescaler {
  std::string err;
  cudaStream_t stream = nullptr;
  std::vector<void*> allocs;
  int64_t m = 0, n = 0, nnz = 0, qnnz = 0;
  // A: CSC + column of every entry + CSR view (rowptr, perm)
  int *colptr = nullptr, *rowval = nullptr, *colof = nullptr, *rowptr = nullptr, *perm = nullptr;
  double* val = nullptr;
  // Q: CSC
  int *qcolptr = nullptr, *qrowval = nullptr, *qcolof = nullptr;
  double* qval = nullptr;
  double *c = nullptr, *l = nullptr, *u = nullptr, *b = nullptr;
  double *con = nullptr, *var = nullptr, *cum_con = nullptr, *cum_var = nullptr, *ruiz_con = nullptr,
         *ruiz_var = nullptr, *t1 = nullptr, *t2 = nullptr, *sf = nullptr;
  int *long_cols = nullptr, *long_rows = nullptr, *long_qcols = nullptr;
  int n_long_cols = 0, n_long_rows = 0, n_long_qcols = 0;
  int64_t launches = 0;

  ~Rescaler() {
    for (void* p : allocs) cudaFree(p);
    if (stream) cudaStreamDestroy(stream);
  }
  template <class T>
  bool alloc(T** p, int64_t count) {
    void* q = nullptr;
    const cudaError_t e = cudaMalloc(&q, static_cast<size_t>(std::max<int64_t>(count, 1)) * sizeof(T));
    if (e != cudaSuccess) {
      err = std::string("cudaMalloc: ") + cudaGetErrorString(e);
      return false;
    }
    allocs.push_back(q);
    *p = static_cast<T*>(q);
    return true;
  }
  static int grid_for(int64_t len) {
    return static_cast<int>(std::max<int64_t>(1, std::min<int64_t>((len + kThreads - 1) / kThreads, 148 * 16)));
  }
  void vec(int op, double* a, const double* b2, int64_t len, double s = 0.0, double t = 0.0,
           const int* ptr = nullptr) {
    if (len <= 0) return;
    k_vec<<<grid_for(len), kThreads, 0, stream>>>(op, a, b2, len, s, t, ptr);
    launches += 1;
  }
  // out[s] = statistic of segment s; by columns of A (dim 1), rows of A (dim 2) or columns of Q (dim 3)
  void seg_stat(int dim, int stat, const double* sf_in, double e, double* out) {
    const int nseg = static_cast<int>(dim == 2 ? m : n);
    if (nseg == 0) return;
    const int blocks = (nseg + kThreads - 1) / kThreads;
    if (dim == 1) {
      k_seg_stat<false><<<blocks, kThreads, 0, stream>>>(colptr, nullptr, val, nseg, stat, sf_in, e, out);
      if (n_long_cols)
        k_seg_stat_long<false><<<n_long_cols, kThreads, 0, stream>>>(colptr, nullptr, val, long_cols, stat, sf_in, e, out);
    } else if (dim == 2) {
      k_seg_stat<true><<<blocks, kThreads, 0, stream>>>(rowptr, perm, val, nseg, stat, sf_in, e, out);
      if (n_long_rows)
        k_seg_stat_long<true><<<n_long_rows, kThreads, 0, stream>>>(rowptr, perm, val, long_rows, stat, sf_in, e, out);
    } else {
      k_seg_stat<false><<<blocks, kThreads, 0, stream>>>(qcolptr, nullptr, qval, nseg, stat, sf_in, e, out);
      if (n_long_qcols)
        k_seg_stat_long<false><<<n_long_qcols, kThreads, 0, stream>>>(qcolptr, nullptr, qval, long_qcols, stat, sf_in, e, out);
    }
    launches += 2;
  }
  // l2_norm, preprocess.jl:99-113: scale_factor = max |.| (0 -> 1); out = scale_factor * sqrt(sum (v / scale_factor)^2)
  void l2_norm(int dim, double* out) {
    const int64_t len = dim == 2 ? m : n;
    seg_stat(dim, kMaxAbs, nullptr, 0.0, sf);
    vec(kZeroToOne, sf, nullptr, len);
    seg_stat(dim, kSumSqScaled, sf, 0.0, out);
    vec(kL2Finish, out, sf, len);
  }
  // scale_problem, preprocess.jl:555-573, with con / var
  void scale_problem() {
    vec(kDivBy, c, var, n);
    if (qnnz) {
      k_scale_entries<<<grid_for(qnnz), kThreads, 0, stream>>>(qrowval, qcolof, qnnz, var, var, qval);
      launches += 1;
    }
    vec(kMulBy, u, var, n);
    vec(kMulBy, l, var, n);
    vec(kDivBy, b, con, m);
    if (nnz) {
      k_scale_entries<<<grid_for(nnz), kThreads, 0, stream>>>(rowval, colof, nnz, con, var, val);
      launches += 1;
    }
  }
};

}  // namespace

// rescale_problem (src/preprocess.jl:631-687) computed on the current CUDA device. Same contract as
// the reference function: the arrays of the (copied) problem are rescaled IN PLACE and the
// cumulative constraint / variable rescaling vectors are returned, so that
// ScaledQpProblem(original, scaled, constraint_rescaling, variable_rescaling) can be formed.
extern "C" int folp_rescale_problem(int64_t m, int64_t n, int32_t index_base, const int64_t* colptr,
                                    const int64_t* rowval, double* nzval, const int64_t* q_colptr,
                                    const int64_t* q_rowval, double* q_nzval, double* c, double* l,
                                    double* u, double* b, int32_t l_inf_ruiz_iterations, int32_t ruiz_p,
                                    int32_t l2_norm_rescaling, double pock_chambolle_alpha,
                                    double* constraint_rescaling, double* variable_rescaling) {
  Rescaler R;
  auto fail = [&](int code, const std::string& msg) {
    set_create_error(msg);
    return code;
  };
  if (m < 0 || n < 0 || !colptr || (index_base != 0 && index_base != 1) || !constraint_rescaling ||
      !variable_rescaling || (n > 0 && (!c || !l || !u)) || (m > 0 && !b))
    return fail(FOLP_INVALID_ARGUMENT, "folp_rescale_problem: invalid argument");
  if (ruiz_p != 0 && ruiz_p != 2)
    return fail(FOLP_INVALID_ARGUMENT, "folp_rescale_problem: ruiz_p must be 0 (infinity norm) or 2");
  const int base = index_base;
  const int64_t nnz = colptr[n] - base;
  const int64_t qnnz = q_colptr ? q_colptr[n] - base : 0;
  if (nnz < 0 || qnnz < 0 || (nnz > 0 && (!rowval || !nzval)) || (qnnz > 0 && (!q_rowval || !q_nzval)))
    return fail(FOLP_INVALID_ARGUMENT, "folp_rescale_problem: invalid matrix arrays");
  if (n + m >= (int64_t{1} << 31) - 64 || nnz >= (int64_t{1} << 31) - 64 || qnnz >= (int64_t{1} << 31) - 64)
    return fail(FOLP_UNSUPPORTED, "folp_rescale_problem: problem exceeds 32-bit indexing");
  int device = 0;
  cudaDeviceProp prop;
  if (cudaGetDevice(&device) != cudaSuccess || cudaGetDeviceProperties(&prop, device) != cudaSuccess)
    return fail(FOLP_CUDA_ERROR, "folp_rescale_problem: no CUDA device");
  if (prop.major != 10)
    return fail(FOLP_UNSUPPORTED, std::string("libfolp_b200 is built for sm_100a only; device is ") + prop.name);
  R.m = m; R.n = n; R.nnz = nnz; R.qnnz = qnnz;

  // ---- host: CSR view of A as a permutation of the CSC positions; long rows / columns ----
  std::vector<int> h_rowptr, h_perm;
  if (!csr_permutation(n, m, nnz, colptr, rowval, base, &h_rowptr, &h_perm))
    return fail(FOLP_INVALID_ARGUMENT, "folp_rescale_problem: row index out of range");
  std::vector<int> h_long_cols, h_long_rows, h_long_qcols;
  for (int64_t j = 0; j < n; ++j) {
    if (colptr[j + 1] - colptr[j] > kLongSegment) h_long_cols.push_back(static_cast<int>(j));
    if (qnnz && q_colptr[j + 1] - q_colptr[j] > kLongSegment) h_long_qcols.push_back(static_cast<int>(j));
  }
  for (int64_t i = 0; i < m; ++i)
    if (h_rowptr[i + 1] - h_rowptr[i] > kLongSegment) h_long_rows.push_back(static_cast<int>(i));

#define RS_TRY(expr)                                                              \
  do {                                                                            \
    cudaError_t _e = (expr);                                                      \
    if (_e != cudaSuccess)                                                        \
      return fail(_e == cudaErrorMemoryAllocation ? FOLP_OUT_OF_MEMORY : FOLP_CUDA_ERROR, \
                  std::string(#expr) + ": " + cudaGetErrorString(_e));            \
  } while (0)
#define RS_ALLOC(p, count) \
  do { if (!R.alloc(&(p), (count))) return fail(FOLP_OUT_OF_MEMORY, R.err); } while (0)

  RS_TRY(cudaStreamCreateWithFlags(&R.stream, cudaStreamNonBlocking));
  cudaStream_t s = R.stream;
  // ---- upload; Int64 1-based indices are narrowed on the device ----
  int64_t* staging = nullptr;  // Int64 staging, reused
  RS_ALLOC(staging, std::max<int64_t>(std::max(nnz, qnnz), n + 1));
  RS_ALLOC(R.colptr, n + 1);
  RS_ALLOC(R.rowval, nnz);
  RS_ALLOC(R.colof, nnz);
  RS_ALLOC(R.val, nnz);
  RS_ALLOC(R.rowptr, m + 1);
  RS_ALLOC(R.perm, nnz);
  auto narrow = [&](const int64_t* src, int64_t len, int* dst) -> cudaError_t {
    if (len <= 0) return cudaSuccess;
    cudaError_t e = cudaMemcpyAsync(staging, src, sizeof(int64_t) * len, cudaMemcpyHostToDevice, s);
    if (e != cudaSuccess) return e;
    k_narrow<<<Rescaler::grid_for(len), kThreads, 0, s>>>(staging, base, len, dst);
    return cudaGetLastError();
  };
  RS_TRY(narrow(colptr, n + 1, R.colptr));
  RS_TRY(narrow(rowval, nnz, R.rowval));
  if (nnz) {
    RS_TRY(cudaMemcpyAsync(R.val, nzval, sizeof(double) * nnz, cudaMemcpyHostToDevice, s));
    RS_TRY(cudaMemcpyAsync(R.perm, h_perm.data(), sizeof(int) * nnz, cudaMemcpyHostToDevice, s));
    k_col_of<<<Rescaler::grid_for(nnz), kThreads, 0, s>>>(R.colptr, static_cast<int>(n), static_cast<int>(nnz), R.colof);
  }
  RS_TRY(cudaMemcpyAsync(R.rowptr, h_rowptr.data(), sizeof(int) * (m + 1), cudaMemcpyHostToDevice, s));
  if (qnnz) {
    RS_ALLOC(R.qcolptr, n + 1);
    RS_ALLOC(R.qrowval, qnnz);
    RS_ALLOC(R.qcolof, qnnz);
    RS_ALLOC(R.qval, qnnz);
    RS_TRY(narrow(q_colptr, n + 1, R.qcolptr));
    RS_TRY(narrow(q_rowval, qnnz, R.qrowval));
    RS_TRY(cudaMemcpyAsync(R.qval, q_nzval, sizeof(double) * qnnz, cudaMemcpyHostToDevice, s));
    k_col_of<<<Rescaler::grid_for(qnnz), kThreads, 0, s>>>(R.qcolptr, static_cast<int>(n), static_cast<int>(qnnz), R.qcolof);
  }
  auto upload_ids = [&](const std::vector<int>& ids, int** d, int* count) -> bool {
    *count = static_cast<int>(ids.size());
    if (ids.empty()) return true;
    if (!R.alloc(d, static_cast<int64_t>(ids.size()))) return false;
    return cudaMemcpyAsync(*d, ids.data(), sizeof(int) * ids.size(), cudaMemcpyHostToDevice, s) == cudaSuccess;
  };
  if (!upload_ids(h_long_cols, &R.long_cols, &R.n_long_cols) ||
      !upload_ids(h_long_rows, &R.long_rows, &R.n_long_rows) ||
      !upload_ids(h_long_qcols, &R.long_qcols, &R.n_long_qcols))
    return fail(FOLP_OUT_OF_MEMORY, "folp_rescale_problem: device allocation failed");
  auto up_vec = [&](double** d, const double* h, int64_t len) -> cudaError_t {
    if (!R.alloc(d, len)) return cudaErrorMemoryAllocation;
    return len ? cudaMemcpyAsync(*d, h, sizeof(double) * len, cudaMemcpyHostToDevice, s) : cudaSuccess;
  };
  RS_TRY(up_vec(&R.c, c, n));
  RS_TRY(up_vec(&R.l, l, n));
  RS_TRY(up_vec(&R.u, u, n));
  RS_TRY(up_vec(&R.b, b, m));
  RS_ALLOC(R.con, m); RS_ALLOC(R.cum_con, m); RS_ALLOC(R.ruiz_con, m);
  RS_ALLOC(R.var, n); RS_ALLOC(R.cum_var, n); RS_ALLOC(R.ruiz_var, n);
  RS_ALLOC(R.t1, n); RS_ALLOC(R.t2, n);
  RS_ALLOC(R.sf, std::max(m, n));
  R.vec(kFill, R.cum_con, nullptr, m, 1.0);
  R.vec(kFill, R.cum_var, nullptr, n, 1.0);

  // ---- ruiz_rescaling, preprocess.jl:412-477 ----
  if (l_inf_ruiz_iterations > 0) {
    R.vec(kFill, R.ruiz_con, nullptr, m, 1.0);
    R.vec(kFill, R.ruiz_var, nullptr, n, 1.0);
    // is_linear_programming_problem: iszero(objective_matrix); scaling never zeroes an entry
    bool is_lp = true;
    for (int64_t k = 0; k < qnnz && is_lp; ++k)
      if (q_nzval[k] != 0.0) is_lp = false;
    for (int it = 0; it < l_inf_ruiz_iterations; ++it) {
      if (ruiz_p == 0) {
        R.seg_stat(1, kMaxAbs, nullptr, 0.0, R.var);
        if (qnnz) R.seg_stat(3, kMaxAbs, nullptr, 0.0, R.t2);
        else R.vec(kFill, R.t2, nullptr, n, 0.0);
        R.vec(kSqrtMax, R.var, R.t2, n);
      } else {
        R.l2_norm(1, R.var);
        if (qnnz) R.l2_norm(3, R.t2);
        else R.vec(kFill, R.t2, nullptr, n, 0.0);
        R.vec(kSqrtSqrtSumSq, R.var, R.t2, n);
      }
      if (m > 0) {
        if (ruiz_p == 0) {
          R.seg_stat(2, kMaxAbs, nullptr, 0.0, R.con);
          R.vec(kSqrt, R.con, nullptr, m);
        } else {
          R.l2_norm(2, R.con);
          const double target = is_lp ? sqrt(static_cast<double>(n) / static_cast<double>(m))
                                      : sqrt(static_cast<double>(n) / static_cast<double>(m + n));
          R.vec(kSqrtDiv, R.con, nullptr, m, target);
        }
      }
      R.scale_problem();
      R.vec(kMulBy, R.ruiz_con, R.con, m);
      R.vec(kMulBy, R.ruiz_var, R.var, n);
    }
    R.vec(kMulBy, R.cum_con, R.ruiz_con, m);  // preprocess.jl:650-651
    R.vec(kMulBy, R.cum_var, R.ruiz_var, n);
  }
  // ---- l2_norm_rescaling, preprocess.jl:358-372 ----
  if (l2_norm_rescaling) {
    R.l2_norm(2, R.con);
    R.l2_norm(1, R.var);
    R.vec(kZeroToOneSqrt, R.con, nullptr, m);
    R.vec(kZeroToOneSqrt, R.var, nullptr, n);
    R.scale_problem();
    R.vec(kMulBy, R.cum_con, R.con, m);
    R.vec(kMulBy, R.cum_var, R.var, n);
  }
  // ---- pock_chambolle_rescaling, preprocess.jl:508-539 ----
  if (pock_chambolle_alpha >= 0.0) {
    const double alpha = pock_chambolle_alpha;
    // mapreduce over a sparse matrix also folds f(0) for every structural zero
    const double zero_col = pow(0.0, 2 - alpha), zero_row = pow(0.0, alpha);
    R.seg_stat(1, kSumPow, nullptr, 2 - alpha, R.var);
    R.seg_stat(2, kSumPow, nullptr, alpha, R.con);
    R.vec(kAddCountSqrt, R.var, nullptr, n, zero_col, static_cast<double>(m), R.colptr);
    R.vec(kAddCountSqrt, R.con, nullptr, m, zero_row, static_cast<double>(n), R.rowptr);
    R.scale_problem();
    R.vec(kMulBy, R.cum_con, R.con, m);
    R.vec(kMulBy, R.cum_var, R.var, n);
  }
  RS_TRY(cudaGetLastError());
  // ---- download ----
  if (nnz) RS_TRY(cudaMemcpyAsync(nzval, R.val, sizeof(double) * nnz, cudaMemcpyDeviceToHost, s));
  if (qnnz) RS_TRY(cudaMemcpyAsync(q_nzval, R.qval, sizeof(double) * qnnz, cudaMemcpyDeviceToHost, s));
  if (n) {
    RS_TRY(cudaMemcpyAsync(c, R.c, sizeof(double) * n, cudaMemcpyDeviceToHost, s));
    RS_TRY(cudaMemcpyAsync(l, R.l, sizeof(double) * n, cudaMemcpyDeviceToHost, s));
    RS_TRY(cudaMemcpyAsync(u, R.u, sizeof(double) * n, cudaMemcpyDeviceToHost, s));
    RS_TRY(cudaMemcpyAsync(variable_rescaling, R.cum_var, sizeof(double) * n, cudaMemcpyDeviceToHost, s));
  }
  if (m) {
    RS_TRY(cudaMemcpyAsync(b, R.b, sizeof(double) * m, cudaMemcpyDeviceToHost, s));
    RS_TRY(cudaMemcpyAsync(constraint_rescaling, R.cum_con, sizeof(double) * m, cudaMemcpyDeviceToHost, s));
  }
  RS_TRY(cudaStreamSynchronize(s));
  if (getenv("FOLP_TIMING")) fprintf(stderr, "[folp_rescale_problem] %lld kernel launches\n", static_cast<long long>(R.launches));
  return FOLP_OK;
#undef RS_TRY
#undef RS_ALLOC
}
