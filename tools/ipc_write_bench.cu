// ipc_write_bench.cu -- development probe: does exporting a cudaMalloc allocation through CUDA IPC
// (cudaIpcGetMemHandle) change how fast the OWNING GPU streams into it / gathers from it?
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

__global__ void k_write(double2* __restrict__ dst, const double2* __restrict__ src, int n2) {
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n2; j += gridDim.x * blockDim.x) {
    double2 v = src[j];
    v.x += 1.0;
    dst[j] = v;
  }
}
__global__ void k_gather(const double* __restrict__ tab, const int* __restrict__ idx, int n, double* out) {
  double acc = 0.0;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) acc += __ldg(tab + idx[j]);
  if (acc == 1.2345) out[0] = acc;
}
static float time_it(void (*launch)(void*), void* ctx) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int i = 0; i < 5; ++i) launch(ctx);
  CK(cudaEventRecord(e0));
  for (int i = 0; i < 50; ++i) launch(ctx);
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  return ms / 50 * 1e3f;
}
struct Ctx { double2* dst; const double2* src; int n2; const double* tab; const int* idx; int n; double* out; };
static void l_write(void* p) { Ctx* c = (Ctx*)p; k_write<<<592, 256>>>(c->dst, c->src, c->n2); }
static void l_gather(void* p) { Ctx* c = (Ctx*)p; k_gather<<<592, 256>>>(c->tab, c->idx, c->n, c->out); }

int main() {
  const int n = 1000000, ng = 5000000;
  double *a, *b, *out; int* idx;
  CK(cudaMalloc(&a, n * 8)); CK(cudaMalloc(&b, n * 8)); CK(cudaMalloc(&out, 8)); CK(cudaMalloc(&idx, ng * 4));
  CK(cudaMemset(a, 0, n * 8)); CK(cudaMemset(b, 0, n * 8));
  int* h = (int*)malloc(ng * 4);
  unsigned long long s = 88172645463325252ull;
  for (int k = 0; k < ng; ++k) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; h[k] = (int)(s % n); }
  CK(cudaMemcpy(idx, h, ng * 4, cudaMemcpyHostToDevice));
  Ctx c{(double2*)b, (const double2*)a, n / 4, b, idx, ng, out};
  printf("before export: write 4 MB %.2f us, 5e6 gathers over 8 MB %.2f us\n", time_it(l_write, &c), time_it(l_gather, &c));
  cudaIpcMemHandle_t hd;
  CK(cudaIpcGetMemHandle(&hd, b));
  printf("after  export: write 4 MB %.2f us, 5e6 gathers over 8 MB %.2f us\n", time_it(l_write, &c), time_it(l_gather, &c));
  return 0;
}
