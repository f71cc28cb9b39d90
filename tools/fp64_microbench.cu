// FP64 micro-benchmarks for B200 (sm_100a): DFMA latency / throughput, pow() latency / throughput.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_microbench fp64_microbench.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void dfma_latency(double* out, long long* cyc, int iters, double a, double b) {
  double x = threadIdx.x * 1e-3;
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) x = fma(x, a, b);
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = x;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int ILP>
__global__ void dfma_tput(double* out, int iters, double a, double b) {
  double x[ILP];
#pragma unroll
  for (int k = 0; k < ILP; k++) x[k] = threadIdx.x * 1e-3 + k;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < ILP; k++) x[k] = fma(x[k], a, b);
  }
  double s = 0;
#pragma unroll
  for (int k = 0; k < ILP; k++) s += x[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void pow_latency(double* out, long long* cyc, int iters, double e) {
  double x = 0.5 + threadIdx.x * 1e-3;
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) x = pow(x, e) + 0.25;
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = x;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int ILP>
__global__ void pow_tput(double* out, int iters, double e) {
  double x[ILP];
#pragma unroll
  for (int k = 0; k < ILP; k++) x[k] = 0.5 + threadIdx.x * 1e-3 + 0.01 * k;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < ILP; k++) x[k] = pow(x[k], e) + 0.25;
  }
  double s = 0;
#pragma unroll
  for (int k = 0; k < ILP; k++) s += x[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void ddiv_sqrt_latency(double* out, long long* cyc, int iters, double d) {
  double x = 1.5 + threadIdx.x * 1e-3, y = x;
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) x = x / d + 1.0;
  long long t1 = clock64();
  for (int i = 0; i < iters; i++) y = sqrt(y) + 1.0;
  long long t2 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = x + y;
  if (threadIdx.x == 0 && blockIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = t2 - t1; }
}

template <class F>
float time_ms(F f) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int sms = p.multiProcessorCount;
  printf("%s SMs=%d clock=%d kHz\n", p.name, sms, p.clockRate);
  double* out; cudaMalloc(&out, sizeof(double) * sms * 32 * 1024);
  long long* cyc; cudaMallocManaged(&cyc, 16);
  const int it = 20000;
  dfma_latency<<<1, 32>>>(out, cyc, it, 0.999, 1e-3); cudaDeviceSynchronize();
  dfma_latency<<<1, 32>>>(out, cyc, it, 0.999, 1e-3); cudaDeviceSynchronize();
  printf("DFMA dependent latency: %.2f cycles\n", (double)*cyc / it);
  pow_latency<<<1, 32>>>(out, cyc, 2000, 1.37); cudaDeviceSynchronize();
  pow_latency<<<1, 32>>>(out, cyc, 2000, 1.37); cudaDeviceSynchronize();
  printf("pow dependent latency (1 warp): %.1f cycles\n", (double)*cyc / 2000);
  ddiv_sqrt_latency<<<1, 32>>>(out, cyc, 2000, 1.0000001); cudaDeviceSynchronize();
  printf("div latency %.1f cycles, sqrt latency %.1f cycles\n", (double)cyc[0] / 2000, (double)cyc[1] / 2000);
  // throughput: DFMA
  for (int wpsm : {4, 8, 16, 32, 64}) {
    int threads = 256, blocks = sms * wpsm * 32 / threads; if (blocks < sms) { threads = wpsm * 32; blocks = sms; }
    float ms1 = time_ms([&] { dfma_tput<1><<<blocks, threads>>>(out, it, 0.999, 1e-3); });
    float ms4 = time_ms([&] { dfma_tput<4><<<blocks, threads>>>(out, it, 0.999, 1e-3); });
    double n = (double)blocks * threads * it;
    printf("DFMA warps/SM=%2d: ILP1 %.2f TFLOP/s, ILP4 %.2f TFLOP/s\n", wpsm, 2 * n / ms1 / 1e9, 2 * 4 * n / ms4 / 1e9);
  }
  for (int wpsm : {4, 8, 12, 16, 32, 64}) {
    int threads = 128, blocks = sms * wpsm * 32 / threads;
    const int pit = 2000;
    float ms1 = time_ms([&] { pow_tput<1><<<blocks, threads>>>(out, pit, 1.37); });
    float ms2 = time_ms([&] { pow_tput<2><<<blocks, threads>>>(out, pit, 1.37); });
    float ms4 = time_ms([&] { pow_tput<4><<<blocks, threads>>>(out, pit, 1.37); });
    double n = (double)blocks * threads * pit;
    printf("pow  warps/SM=%2d: ILP1 %.1f Gpow/s, ILP2 %.1f Gpow/s, ILP4 %.1f Gpow/s\n", wpsm, n / ms1 / 1e6, 2 * n / ms2 / 1e6,
           4 * n / ms4 / 1e6);
  }
  return 0;
}
