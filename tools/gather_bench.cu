// gather_bench.cu -- development microbenchmark: how many random 8-byte gathers per clock
// per SM does a B200 sustain from an L2-resident (or larger) table, by load flavour, threads
// per CTA, independent loads per thread and shared-memory carve-out?
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o gather_bench gather_bench.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#define CK(x)                                                                   \
  do {                                                                          \
    cudaError_t e = (x);                                                        \
    if (e != cudaSuccess) {                                                     \
      printf("%s: %s\n", #x, cudaGetErrorString(e));                            \
      exit(1);                                                                  \
    }                                                                           \
  } while (0)

template <int MODE>
__device__ __forceinline__ double ld(const double* p) {
  double v;
  if (MODE == 0) asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(v) : "l"(p));
  else if (MODE == 1) asm volatile("ld.global.cg.f64 %0, [%1];" : "=d"(v) : "l"(p));
  else if (MODE == 2) asm volatile("ld.global.ca.f64 %0, [%1];" : "=d"(v) : "l"(p));
  else if (MODE == 3) asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
  else asm volatile("ld.global.cv.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}

template <int MODE, int U>
__global__ void k_gather(const int* __restrict__ idx, const double* __restrict__ tab, long long nnz,
                         double* out) {
  extern __shared__ unsigned char smem[];
  const long long stride = (long long)gridDim.x * blockDim.x * U;
  double acc = 0.0;
  for (long long base = (long long)blockIdx.x * blockDim.x * U + threadIdx.x; base < nnz;
       base += stride) {
    int c[U];
    double v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long k = base + (long long)u * blockDim.x;
      c[u] = k < nnz ? idx[k] : -1;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = c[u] >= 0 ? ld<MODE>(tab + c[u]) : 0.0;
#pragma unroll
    for (int u = 0; u < U; ++u) acc += v[u];
  }
  if (acc == 123.456) out[0] = acc + smem[0];
}

// The same gathers with the matrix stream beside them: acc += val[k] * tab[idx[k]], i.e. a CSR SpMV
// without any row structure -- the ceiling of every SpMV design for this access pattern.
template <int U>
__global__ void k_stream_gather(const int* __restrict__ idx, const double* __restrict__ val,
                                const double* __restrict__ tab, long long nnz, double* out) {
  const long long stride = (long long)gridDim.x * blockDim.x * U;
  double acc = 0.0;
  for (long long base = (long long)blockIdx.x * blockDim.x * U + threadIdx.x; base < nnz;
       base += stride) {
    int c[U];
    double a[U], v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long k = base + (long long)u * blockDim.x;
      c[u] = k < nnz ? idx[k] : -1;
      a[u] = k < nnz ? val[k] : 0.0;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = c[u] >= 0 ? ld<0>(tab + c[u]) : 0.0;
#pragma unroll
    for (int u = 0; u < U; ++u) acc += a[u] * v[u];
  }
  if (acc == 123.456) out[0] = acc;
}
template <int U>
void run_stream(const int* idx, const double* val, const double* tab, long long nnz, double* out,
                int threads, int ctas_per_sm) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  const int grid = 148 * ctas_per_sm;
  for (int w = 0; w < 3; ++w) k_stream_gather<U><<<grid, threads>>>(idx, val, tab, nnz, out);
  CK(cudaEventRecord(e0));
  const int reps = 10;
  for (int r = 0; r < reps; ++r) k_stream_gather<U><<<grid, threads>>>(idx, val, tab, nnz, out);
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  float ms;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  const double us = 1e3 * ms / reps;
  printf("stream+gather  U=%2d thr=%4d cta/sm=%d              : %8.2f us  %6.1f Gnnz/s  %.3f /clk/SM@1.965  (12 B/nnz stream = %.0f GB/s)\n",
         U, threads, ctas_per_sm, us, nnz / us * 1e-3, nnz / us * 1e-3 / (148 * 1.965), 12.0 * nnz / us * 1e-3);
}

template <int MODE, int U>
void run(const char* name, const int* idx, const double* tab, long long nnz, double* out,
         int threads, int ctas_per_sm, int smem) {
  CK(cudaFuncSetAttribute(k_gather<MODE, U>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  const int grid = 148 * ctas_per_sm;
  for (int w = 0; w < 3; ++w) k_gather<MODE, U><<<grid, threads, smem>>>(idx, tab, nnz, out);
  CK(cudaEventRecord(e0));
  const int reps = 10;
  for (int r = 0; r < reps; ++r) k_gather<MODE, U><<<grid, threads, smem>>>(idx, tab, nnz, out);
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  float ms;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  const double us = 1e3 * ms / reps;
  printf("%-14s U=%2d thr=%4d cta/sm=%d smem=%6d : %8.2f us  %6.1f Ggather/s  %.3f /clk/SM@1.965\n",
         name, U, threads, ctas_per_sm, smem, us, nnz / us * 1e-3, nnz / us * 1e-3 / (148 * 1.965));
}

int main(int argc, char** argv) {
  const long long n = argc > 1 ? atoll(argv[1]) : 1000000;
  const long long nnz = argc > 2 ? atoll(argv[2]) : 10000000;
  std::vector<int> h(nnz);
  uint64_t s = 88172645463325252ull;
  for (long long k = 0; k < nnz; ++k) {
    s ^= s << 13; s ^= s >> 7; s ^= s << 17;
    h[k] = (int)(s % (uint64_t)n);
  }
  int* idx;
  double *tab, *out;
  CK(cudaMalloc(&idx, nnz * 4));
  CK(cudaMalloc(&tab, n * 8));
  CK(cudaMalloc(&out, 8));
  CK(cudaMemcpy(idx, h.data(), nnz * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(tab, 0, n * 8));
  double* val;
  CK(cudaMalloc(&val, nnz * 8));
  CK(cudaMemset(val, 0, nnz * 8));
  printf("table %lld doubles (%.1f MB), %lld gathers\n", n, n * 8e-6, nnz);
  run_stream<4>(idx, val, tab, nnz, out, 512, 1);
  run_stream<4>(idx, val, tab, nnz, out, 1024, 1);
  run_stream<4>(idx, val, tab, nnz, out, 1024, 2);
  run_stream<8>(idx, val, tab, nnz, out, 1024, 2);
  run_stream<2>(idx, val, tab, nnz, out, 1024, 2);
  const int big = 160 * 1024;
  run<0, 4>("nc", idx, tab, nnz, out, 512, 1, big);
  run<0, 8>("nc", idx, tab, nnz, out, 512, 1, big);
  run<0, 4>("nc", idx, tab, nnz, out, 1024, 1, big);
  run<0, 8>("nc", idx, tab, nnz, out, 1024, 1, big);
  run<0, 4>("nc", idx, tab, nnz, out, 1024, 2, 0);
  run<0, 8>("nc", idx, tab, nnz, out, 1024, 2, 0);
  run<1, 4>("cg", idx, tab, nnz, out, 512, 1, big);
  run<1, 8>("cg", idx, tab, nnz, out, 1024, 1, big);
  run<1, 8>("cg", idx, tab, nnz, out, 1024, 2, 0);
  run<2, 8>("ca", idx, tab, nnz, out, 1024, 1, big);
  run<2, 8>("ca", idx, tab, nnz, out, 1024, 2, 0);
  run<3, 4>("nc.noalloc", idx, tab, nnz, out, 512, 1, big);
  run<3, 8>("nc.noalloc", idx, tab, nnz, out, 1024, 1, big);
  run<3, 8>("nc.noalloc", idx, tab, nnz, out, 1024, 2, 0);
  run<4, 8>("cv", idx, tab, nnz, out, 1024, 1, big);
  return 0;
}
