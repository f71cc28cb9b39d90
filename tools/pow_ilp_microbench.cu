// Does interleaving N independent pow chains pay on B200?  Times pow_core_v<N> (lgar_pow.cuh) with 1, 2, 4 warps
// per SM sub-partition.  Build twice: ptxas -O3 re-serialises the chains, -O1 keeps the interleaved source order.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 --fmad=false -std=c++17 [-Xptxas -O1] -o pow_ilp tools/pow_ilp_microbench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../lgar-py_b200/csrc/lgar_pow.cuh"
using namespace lgar;

template <int N>
__global__ void __launch_bounds__(512, 1) bench(double* out, int iters, double y) {
  pow_tables_to_shared();
  double x[N], yv[N], r[N];
  bool ok[N];
  for (int k = 0; k < N; k++) { x[k] = 0.5 + threadIdx.x * 1e-3 + 0.01 * k; yv[k] = y; }
  for (int i = 0; i < iters; i++) {
    pow_core_v<N>(x, yv, r, ok);
#pragma unroll
    for (int k = 0; k < N; k++) x[k] = r[k] + 0.25;  // next input depends on this output
  }
  double s = 0;
  for (int k = 0; k < N; k++) s += x[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int N>
void run(int threads, double* d) {
  const int iters = 20000;
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  bench<N><<<148, threads>>>(d, 100, 0.37);
  cudaEventRecord(a);
  bench<N><<<148, threads>>>(d, iters, 0.37);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  const double cyc = ms * 1e-3 * 1.965e9 / iters;
  printf("N=%d warps/SMSP=%d: %.1f cycles per call, %.1f cycles per pow per warp, %.2f pows/cycle/SM\n", N, threads / 128,
         cyc, cyc / N, (threads / 32.0) * N * 32 / cyc);
}
int main() {
  double* d; cudaMalloc(&d, 148 * 512 * 8);
  for (int th : {128, 256, 512}) { run<1>(th, d); run<2>(th, d); run<4>(th, d); }
  return 0;
}
