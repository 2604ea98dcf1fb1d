// fp64_probe.cu -- measures DFMA latency (dependent chain, 1 warp) and throughput (many
// independent chains, full chip) plus f64 division / sqrt cost on the current GPU.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void lat(double* out, int n) {
  double a = out[0], b = 1.0000001, c = 1e-9;
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) a = fma(a, b, c);
  long long t1 = clock64();
  out[1] = a;
  if (threadIdx.x == 0) out[2] = (double)(t1 - t0) / n;
}
__global__ void divlat(double* out, int n) {
  double a = out[0] + 3.0, b = 1.0000001;
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) a = a / b + 1e-9;
  long long t1 = clock64();
  out[1] = a;
  if (threadIdx.x == 0) out[3] = (double)(t1 - t0) / n;
}
__global__ void thr(double* out, int n) {
  double a[8];
  for (int k = 0; k < 8; ++k) a[k] = out[0] + k;
  const double b = 1.0000001, c = 1e-9;
  for (int i = 0; i < n; ++i)
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = fma(a[k], b, c);
  double s = 0;
  for (int k = 0; k < 8; ++k) s += a[k];
  if (s == 12345.678) out[1] = s;
}
int main() {
  double* d;
  cudaMalloc(&d, 64);
  cudaMemset(d, 0, 64);
  lat<<<1, 32>>>(d, 100000);
  divlat<<<1, 32>>>(d, 20000);
  cudaDeviceSynchronize();
  double h[4];
  cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost);
  printf("DFMA dependent latency: %.1f cycles; (div + add) dependent: %.1f cycles\n", h[2], h[3]);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  int n = 20000;
  thr<<<148 * 8, 256>>>(d, 100);
  cudaEventRecord(e0);
  thr<<<148 * 8, 256>>>(d, n);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  double flops = 2.0 * 148 * 8 * 256 * (double)n * 8;
  printf("DFMA throughput: %.2f TFLOP/s (%.3f ms)\n", flops / ms / 1e9, ms);
  return 0;
}
