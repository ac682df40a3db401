// mc_probe.cu -- development probe (one process, all GPUs of the box): what does an all-to-all push of
// one block of doubles per GPU cost over NVLink / NVSwitch, (a) as P-1 posted unicast stores per
// element (what k_take_steps does), (b) as ONE multimem.st per element into an NVSwitch multicast
// object (NVLS), each (i) as an address-ordered stream and (ii) as 256-byte groups in scattered order
// (how the dual epilogue finishes its row groups), and each alone or behind an SpMV-like producer
// (10 random 8-byte gathers from a local table + a 12 B/nonzero stream per element)?
//
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o mc_probe tools/mc_probe.cu -lcuda
//   ./mc_probe [elements_per_gpu = 1250000] [table_elements = 10000000] [reps = 20]
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)
#define CU(x) do { CUresult e_ = (x); if (e_ != CUDA_SUCCESS) { const char* s_ = nullptr; cuGetErrorString(e_, &s_); printf("%s: %s\n", #x, s_ ? s_ : "?"); return false; } } while (0)

constexpr int kMaxDev = 8;
constexpr int kNnzPerRow = 10;

struct Args {
  int me, world;
  long long block;            // elements per GPU
  double* local;              // this GPU's region: world * block doubles
  double* peer[kMaxDev];      // unicast views of every GPU's region
  double* mc;                 // multicast view (one store lands in every GPU's region)
  const double* src;          // block doubles (copy producer)
  const int* cols;            // SpMV-like producer: position-major groups of 32 rows
  const double* vals;
  const double* tab;
  long long ntab;
};

__device__ __forceinline__ void mc_store(double* p, double v) {
  asm volatile("multimem.st.relaxed.sys.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}

// MODE 0: local store only, 1: + world-1 unicast stores, 2: one multimem store (covers the local copy as well)
template <int MODE, bool SCATTER, bool SPMV>
__global__ void __launch_bounds__(256, 4) k_push(Args a) {
  const int lane = threadIdx.x & 31;
  const long long groups = a.block / 32;
  const long long nw = static_cast<long long>(gridDim.x) * (blockDim.x >> 5);
  for (long long g = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5); g < groups; g += nw) {
    const long long gg = SCATTER ? (g * 1000003ll + 12345ll) % groups : g;
    const long long i = gg * 32 + lane;
    double v;
    if (SPMV) {
      v = 0.0;
      const long long base = gg * kNnzPerRow * 32 + lane;
#pragma unroll
      for (int k = 0; k < kNnzPerRow; ++k) v += __ldcs(a.vals + base + k * 32) * __ldg(a.tab + __ldcs(a.cols + base + k * 32));
    } else {
      v = a.src[i];
    }
    const long long at = a.me * a.block + i;
    if (MODE == 2) {
      mc_store(a.mc + at, v);
    } else {
      a.local[at] = v;
      if (MODE == 1) {
#pragma unroll
        for (int r = 0; r < kMaxDev; ++r)
          if (r < a.world && r != a.me) a.peer[r][at] = v;
      }
    }
  }
}

__global__ void k_fill(double* p, long long n, double v) {
  for (long long j = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; j < n; j += static_cast<long long>(gridDim.x) * blockDim.x) p[j] = v;
}
__global__ void k_init_spmv(int* cols, double* vals, double* tab, double* src, long long nnz, long long ntab, long long block, int me) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long j = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; j < nnz; j += stride) {
    unsigned long long s = (j + 1) * 0x9E3779B97F4A7C15ull + me;
    s ^= s >> 29; s *= 0xBF58476D1CE4E5B9ull; s ^= s >> 32;
    cols[j] = static_cast<int>(s % static_cast<unsigned long long>(ntab));
    vals[j] = 1.0;
  }
  for (long long j = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; j < ntab; j += stride) tab[j] = 1.0;
  for (long long j = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; j < block; j += stride) src[j] = me * 1000.0 + (j % 977);
}
__global__ void k_check(const double* region, long long block, int world, int spmv, unsigned long long* bad) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long j = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; j < block * world; j += stride) {
    const int r = static_cast<int>(j / block);
    const long long i = j % block;
    const double want = spmv ? static_cast<double>(kNnzPerRow) : r * 1000.0 + (i % 977);
    if (region[j] != want) atomicAdd(bad, 1ull);
  }
}

struct Dev {
  cudaStream_t st;
  cudaEvent_t e0, e1;
  CUmemGenericAllocationHandle h;
  double *src, *vals, *tab;
  int* cols;
  unsigned long long* bad;
  Args a;
};

static int g_n = 0;
static Dev g_d[kMaxDev];

static bool setup_regions(size_t bytes, bool* have_mc) {
  *have_mc = false;
  CU(cuInit(0));
  int mc_ok = 1;
  for (int d = 0; d < g_n; ++d) {
    int v = 0;
    CU(cuDeviceGetAttribute(&v, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, d));
    printf("device %d: multicast supported = %d\n", d, v);
    mc_ok &= v;
  }
  CUmulticastObjectProp mp;
  memset(&mp, 0, sizeof(mp));
  mp.numDevices = g_n;
  mp.handleTypes = CU_MEM_HANDLE_TYPE_NONE;
  size_t gran = 2u << 20, mc_gran = 0;
  CUmemAllocationProp ap;
  memset(&ap, 0, sizeof(ap));
  ap.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  ap.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  ap.location.id = 0;
  CU(cuMemGetAllocationGranularity(&gran, &ap, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED));
  if (mc_ok) {
    mp.size = bytes;
    CU(cuMulticastGetGranularity(&mc_gran, &mp, CU_MULTICAST_GRANULARITY_RECOMMENDED));
    if (mc_gran > gran) gran = mc_gran;
  }
  const size_t size = (bytes + gran - 1) / gran * gran;
  printf("region %zu bytes, granularity %zu (multicast %zu) -> %zu\n", bytes, gran, mc_gran, size);
  std::vector<CUmemAccessDesc> acc(g_n);
  for (int d = 0; d < g_n; ++d) {
    acc[d].location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    acc[d].location.id = d;
    acc[d].flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  }
  CUdeviceptr uc[kMaxDev];
  for (int d = 0; d < g_n; ++d) {
    ap.location.id = d;
    CU(cuMemCreate(&g_d[d].h, size, &ap, 0));
    CU(cuMemAddressReserve(&uc[d], size, gran, 0, 0));
    CU(cuMemMap(uc[d], size, 0, g_d[d].h, 0));
    CU(cuMemSetAccess(uc[d], size, acc.data(), g_n));
  }
  for (int d = 0; d < g_n; ++d) {
    g_d[d].a.local = reinterpret_cast<double*>(uc[d]);
    for (int r = 0; r < g_n; ++r) g_d[d].a.peer[r] = reinterpret_cast<double*>(uc[r]);
    g_d[d].a.mc = nullptr;
  }
  if (!mc_ok) return true;
  CUmemGenericAllocationHandle mc;
  mp.size = size;
  CU(cuMulticastCreate(&mc, &mp));
  for (int d = 0; d < g_n; ++d) CU(cuMulticastAddDevice(mc, d));
  for (int d = 0; d < g_n; ++d) CU(cuMulticastBindMem(mc, 0, g_d[d].h, 0, size, 0));
  CUdeviceptr mcva;
  CU(cuMemAddressReserve(&mcva, size, gran, 0, 0));
  CU(cuMemMap(mcva, size, 0, mc, 0));
  CU(cuMemSetAccess(mcva, size, acc.data(), g_n));
  for (int d = 0; d < g_n; ++d) g_d[d].a.mc = reinterpret_cast<double*>(mcva);
  *have_mc = true;
  return true;
}

template <int MODE, bool SCATTER, bool SPMV>
static void run(const char* name, int reps, int grid) {
  const Args& a0 = g_d[0].a;
  for (int d = 0; d < g_n; ++d) {
    CK(cudaSetDevice(d));
    k_fill<<<592, 256, 0, g_d[d].st>>>(g_d[d].a.local, a0.block * g_n, -1.0);
  }
  for (int d = 0; d < g_n; ++d) { CK(cudaSetDevice(d)); CK(cudaStreamSynchronize(g_d[d].st)); }
  for (int pass = 0; pass < 2; ++pass) {  // pass 0 warms up
    for (int d = 0; d < g_n; ++d) {
      CK(cudaSetDevice(d));
      CK(cudaEventRecord(g_d[d].e0, g_d[d].st));
      for (int k = 0; k < (pass ? reps : 3); ++k) k_push<MODE, SCATTER, SPMV><<<grid, 256, 0, g_d[d].st>>>(g_d[d].a);
      CK(cudaEventRecord(g_d[d].e1, g_d[d].st));
    }
    for (int d = 0; d < g_n; ++d) { CK(cudaSetDevice(d)); CK(cudaStreamSynchronize(g_d[d].st)); }
  }
  float worst = 0.f, sum = 0.f;
  for (int d = 0; d < g_n; ++d) {
    float ms;
    CK(cudaEventElapsedTime(&ms, g_d[d].e0, g_d[d].e1));
    worst = ms > worst ? ms : worst;
    sum += ms;
  }
  unsigned long long bad_total = 0;
  if (MODE != 0) {
    for (int d = 0; d < g_n; ++d) {
      CK(cudaSetDevice(d));
      CK(cudaMemsetAsync(g_d[d].bad, 0, 8, g_d[d].st));
      k_check<<<592, 256, 0, g_d[d].st>>>(g_d[d].a.local, a0.block, g_n, SPMV ? 1 : 0, g_d[d].bad);
      unsigned long long b = 0;
      CK(cudaMemcpyAsync(&b, g_d[d].bad, 8, cudaMemcpyDeviceToHost, g_d[d].st));
      CK(cudaStreamSynchronize(g_d[d].st));
      bad_total += b;
    }
  }
  const double us = worst / reps * 1e3, us_avg = sum / g_n / reps * 1e3;
  const double in_bytes = static_cast<double>(g_n - 1) * a0.block * 8.0;
  printf("%-44s %9.1f us (avg over GPUs %9.1f)  ingress/GPU %7.1f GB/s  wrong entries %llu\n", name, us, us_avg,
         MODE ? in_bytes / us * 1e-3 : 0.0, bad_total);
  fflush(stdout);
}

int main(int argc, char** argv) {
  long long block = argc > 1 ? atoll(argv[1]) : 1250000;
  const long long ntab = argc > 2 ? atoll(argv[2]) : 10000000;
  const int reps = argc > 3 ? atoi(argv[3]) : 20;
  block = block / 32 * 32;
  CK(cudaGetDeviceCount(&g_n));
  if (g_n > kMaxDev) g_n = kMaxDev;
  if (argc > 4 && atoi(argv[4]) < g_n) g_n = atoi(argv[4]);
  printf("%d GPUs, %lld elements (%.1f MB) per GPU, table %lld, reps %d\n", g_n, block, block * 8e-6, ntab, reps);
  for (int d = 0; d < g_n; ++d) {
    CK(cudaSetDevice(d));
    CK(cudaFree(0));
    for (int r = 0; r < g_n; ++r)
      if (r != d) {
        cudaError_t e = cudaDeviceEnablePeerAccess(r, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { printf("no peer access %d -> %d\n", d, r); return 1; }
        cudaGetLastError();
      }
  }
  bool have_mc = false;
  if (!setup_regions(static_cast<size_t>(block) * g_n * 8, &have_mc)) { printf("region setup failed\n"); return 1; }
  const long long nnz = block * kNnzPerRow;
  for (int d = 0; d < g_n; ++d) {
    CK(cudaSetDevice(d));
    Dev& D = g_d[d];
    CK(cudaStreamCreate(&D.st)); CK(cudaEventCreate(&D.e0)); CK(cudaEventCreate(&D.e1));
    CK(cudaMalloc(&D.src, block * 8)); CK(cudaMalloc(&D.vals, nnz * 8)); CK(cudaMalloc(&D.cols, nnz * 4));
    CK(cudaMalloc(&D.tab, ntab * 8)); CK(cudaMalloc(&D.bad, 8));
    k_init_spmv<<<592, 256, 0, D.st>>>(D.cols, D.vals, D.tab, D.src, nnz, ntab, block, d);
    D.a.me = d; D.a.world = g_n; D.a.block = block; D.a.src = D.src; D.a.cols = D.cols; D.a.vals = D.vals; D.a.tab = D.tab; D.a.ntab = ntab;
    CK(cudaStreamSynchronize(D.st));
  }
  const int grid = 592;
  run<0, false, false>("copy, local only", reps, grid);
  run<1, false, false>("copy, unicast stores, address order", reps, grid);
  run<1, true, false>("copy, unicast stores, scattered 256 B groups", reps, grid);
  if (have_mc) {
    run<2, false, false>("copy, multimem.st, address order", reps, grid);
    run<2, true, false>("copy, multimem.st, scattered 256 B groups", reps, grid);
  }
  run<0, true, true>("SpMV-like producer, local only", reps, grid);
  run<1, true, true>("SpMV-like producer, unicast stores", reps, grid);
  if (have_mc) run<2, true, true>("SpMV-like producer, multimem.st", reps, grid);
  run<0, false, true>("SpMV-like producer in order, local only", reps, grid);
  run<1, false, true>("SpMV-like producer in order, unicast stores", reps, grid);
  if (have_mc) run<2, false, true>("SpMV-like producer in order, multimem.st", reps, grid);
  return 0;
}
