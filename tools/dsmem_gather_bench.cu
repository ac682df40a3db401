// dsmem_gather_bench.cu -- development microbenchmark: random 8-byte gathers from a table held in
// the DISTRIBUTED shared memory of a thread-block cluster (each CTA owns a slice of the table, every
// thread reads any slice through ld.shared::cluster), against the same gathers from L2. Question:
// could a cluster-resident panel of the gathered vector beat the ~0.8 gathers/clk/SM an SM gets out
// of L2 (tools/gather_bench.cu)?
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o dsmem_gather_bench dsmem_gather_bench.cu
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

namespace cg = cooperative_groups;

#define CK(x)                                                                   \
  do {                                                                          \
    cudaError_t e = (x);                                                        \
    if (e != cudaSuccess) {                                                     \
      printf("%s: %s\n", #x, cudaGetErrorString(e));                            \
      exit(1);                                                                  \
    }                                                                           \
  } while (0)

// slice = doubles per CTA. idx values are in [0, cluster_size * slice) (MODE 0: any CTA of the cluster;
// MODE 1: remapped to the thread's own CTA = plain shared memory; MODE 2: from the global copy = L2).
template <int MODE, int U>
__global__ void k_dsmem(const int* __restrict__ idx, const double* __restrict__ tab, long long per_cluster,
                        int slice, int reps, double* out) {
  extern __shared__ double sm[];
  cg::cluster_group cluster = cg::this_cluster();
  const int crank = cluster.block_rank();
  const int csize = cluster.num_blocks();
  const int cluster_id = blockIdx.x / csize;
  for (int i = threadIdx.x; i < slice; i += blockDim.x) sm[i] = tab[(long long)crank * slice + i];
  cluster.sync();
  const int* my = idx + (long long)cluster_id * per_cluster;
  const long long stride = (long long)csize * blockDim.x * U;
  double acc = 0.0;
  for (int r = 0; r < reps; ++r) {
    for (long long base = (long long)crank * blockDim.x * U + threadIdx.x; base < per_cluster; base += stride) {
      int c[U];
      double v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long k = base + (long long)u * blockDim.x;
        c[u] = k < per_cluster ? __ldcs(my + k) : -1;
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (c[u] < 0) { v[u] = 0.0; continue; }
        if (MODE == 2) {
          v[u] = __ldg(tab + c[u]);
        } else if (MODE == 1) {
          v[u] = sm[c[u] % slice];
        } else {
          const int owner = c[u] / slice, off = c[u] - owner * slice;
          const double* remote = cluster.map_shared_rank(sm, owner);
          v[u] = remote[off];
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) acc += v[u];
    }
  }
  cluster.sync();  // nobody leaves while its slice may still be read
  if (acc == 123.456) out[0] = acc;
}

template <int MODE, int U>
void run(const char* name, const int* idx, const double* tab, int csize, int threads, int slice,
         long long per_cluster, double* out) {
  const size_t smem = (size_t)slice * sizeof(double);
  CK(cudaFuncSetAttribute(k_dsmem<MODE, U>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (csize > 8) CK(cudaFuncSetAttribute(k_dsmem<MODE, U>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(threads);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = csize;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cfg.gridDim = dim3(csize);
  int nclusters = 0;
  if (cudaOccupancyMaxActiveClusters(&nclusters, k_dsmem<MODE, U>, &cfg) != cudaSuccess || nclusters < 1) {
    printf("%-10s cluster=%2d thr=%4d slice=%6d: not launchable (%s)\n", name, csize, threads, slice,
           cudaGetErrorString(cudaGetLastError()));
    return;
  }
  cfg.gridDim = dim3(nclusters * csize);
  const int reps = 20;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  CK(cudaLaunchKernelEx(&cfg, k_dsmem<MODE, U>, idx, tab, per_cluster, slice, 2, out));
  CK(cudaEventRecord(e0));
  CK(cudaLaunchKernelEx(&cfg, k_dsmem<MODE, U>, idx, tab, per_cluster, slice, reps, out));
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  float ms;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  const double gathers = (double)per_cluster * nclusters * reps;
  const int sms = nclusters * csize;
  printf("%-10s cluster=%2d x %3d clusters (%3d SMs) thr=%4d U=%d slice=%6d doubles: %8.1f us  %.3f gathers/clk/SM@1.965 (incl. table load)\n",
         name, csize, nclusters, sms, threads, U, slice, 1e3 * ms, gathers / (ms * 1e-3) / (sms * 1.965e9));
}

int main() {
  const int slice = 20480;  // 160 KB per CTA
  const int max_csize = 16;
  const long long per_cluster = 1 << 20;
  const int max_clusters = 148;
  std::vector<int> h_idx((size_t)per_cluster * max_clusters);
  uint64_t s = 88172645463325252ull;
  for (auto& v : h_idx) {
    s ^= s << 13; s ^= s >> 7; s ^= s << 17;
    v = (int)(s % ((uint64_t)slice * 8));  // valid for every cluster size >= 8 (and, modulo, for smaller ones)
  }
  int* idx;
  double *tab, *out;
  CK(cudaMalloc(&idx, h_idx.size() * sizeof(int)));
  CK(cudaMemcpy(idx, h_idx.data(), h_idx.size() * sizeof(int), cudaMemcpyHostToDevice));
  CK(cudaMalloc(&tab, (size_t)slice * max_csize * sizeof(double)));
  CK(cudaMemset(tab, 0, (size_t)slice * max_csize * sizeof(double)));
  CK(cudaMalloc(&out, 64));
  for (int threads : {256, 512, 1024}) {
    run<0, 4>("dsmem", idx, tab, 8, threads, slice, per_cluster, out);
    run<0, 8>("dsmem", idx, tab, 8, threads, slice, per_cluster, out);
    run<0, 4>("dsmem", idx, tab, 16, threads, slice, per_cluster, out);
    run<1, 4>("own smem", idx, tab, 8, threads, slice, per_cluster, out);
    run<2, 4>("L2", idx, tab, 8, threads, slice, per_cluster, out);
  }
  run<0, 4>("dsmem", idx, tab, 4, 1024, slice, per_cluster, out);
  run<0, 4>("dsmem", idx, tab, 2, 1024, slice, per_cluster, out);
  return 0;
}
