#include <cstdio>
#include <cuda_runtime.h>
__constant__ double c_w[512];
template <int NW> __global__ void __launch_bounds__(256) probe(double* out, int iters) {
  double acc[24], u[10];
  for (int i = 0; i < 24; ++i) acc[i] = threadIdx.x * 1e-3 + i;
  for (int i = 0; i < 10; ++i) u[i] = 0.999 + i * 1e-6;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int w = 0; w < NW; ++w) {
      // same structure as the stencil: weight w = (a, n): a = w / 6 % 7, n = w % 6; 4 FMAs per weight
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[(w % 6) * 4 + j] = fma(u[j + (w / 6) % 7], c_w[w], acc[(w % 6) * 4 + j]);
    }
    if (it == iters + 1) for (int i = 0; i < 10; ++i) u[i] += acc[i];
  }
  double s = 0; for (int i = 0; i < 24; ++i) s += acc[i];
  if (s == 123456789.0) out[0] = s;
}
template <int NW> void run(const char* name) {
  double* d; cudaMalloc(&d, 64);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int iters = 2000; int grid = 148 * 4;
  float ms = 0;
  for (int r = 0; r < 2; ++r) { cudaEventRecord(e0); probe<NW><<<grid, 256>>>(d, iters); cudaEventRecord(e1); cudaEventSynchronize(e1); }
  cudaEventElapsedTime(&ms, e0, e1);
  double flops = 2.0 * NW * 4 * (double)iters * grid * 256;
  printf("%s NW=%d: %.2f TFLOP/s (%s)\n", name, NW, flops / ms * 1e-9, cudaGetErrorString(cudaGetLastError()));
}
int main() {
  double h[512]; for (int i = 0; i < 512; ++i) h[i] = 1e-3 * (i + 1);
  cudaMemcpyToSymbol(c_w, h, sizeof(h));
  run<42>("const"); run<126>("const"); run<252>("const"); run<294>("const"); run<336>("const"); run<504>("const");
  return 0;
}
