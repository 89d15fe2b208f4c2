// ILP probe: does a second, symmetric cell per thread overlap with the first in the FP64 pipe?
// Same occupancy (5 blocks x 128 threads per SM, capped through dynamic shared memory), W = 1 or 2 cells per thread,
// 64 dependent "met block"-like steps per cell.  Prints cells/s for both; build: nvcc -arch=sm_100a -O3 -fmad=false.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../topoflow_glacier_b200/csrc/tfg_math.cuh"
using namespace tfg::fm;
extern __shared__ double tfg_tabs_decl[];
__device__ __forceinline__ double step(double T, double q, double P, double uz, double h) {
  const double TK = T + 273.15;
  const double rTK = rcp3(TK);
  const double ip0 = exp_tab(-(0.034 * h) * rTK);
  const double e = (q * P) * rcp3(fma(0.378, q, 0.622)) * 0.01;
  const double en = exp_tab(-(17.3 * T) * rcp3(T + 237.3));
  const double RH = e * en * 0.1636661211129296;
  const double L1 = log_tab(e * 0.1636098885816659);
  const double Td = (257.14 * L1) * rcp3(18.678 - L1);
  const double L = log_tab((10.0 - h) * 100.0);
  const double Wp = 1.12 * exp_tab(0.0614 * Td);
  const double es = 6.11 * exp_tab((17.3 * Td) * rcp3(Td + 237.3));
  const double Dh = (uz * 0.16) * rcp3(L * L);
  return (Dh * (T - Td)) + (Dh * (e - RH * es)) * ip0 + Wp + root7(e * 0.1 * rTK);
}
template <int W>
__global__ void __launch_bounds__(128) k(const double* __restrict__ x, double* __restrict__ y, int n, int iters) {
  for (int i = threadIdx.x; i < kTabDoubles; i += blockDim.x) tfg_tabs[i] = (i < 64) ? kExpTab[i] : kLogTab[(i - 64) >> 1][(i - 64) & 1];
  __syncthreads();
  const int i0 = (blockIdx.x * blockDim.x) * W + threadIdx.x;
  double T[W], q[W], P[W], uz[W], h[W];
#pragma unroll
  for (int w = 0; w < W; ++w) { const double* p = x + i0 + w * 128; T[w] = p[0]; q[w] = p[n]; P[w] = p[2 * n]; uz[w] = p[3 * n]; h[w] = p[4 * n]; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int w = 0; w < W; ++w) { const double r = step(T[w], q[w], P[w], uz[w], h[w]); h[w] = 0.5 + 1e-3 * r; T[w] = -5.0 + 1e-2 * r; q[w] = 0.003 + 1e-6 * r; uz[w] = 1.0 + 1e-3 * r; }  // carried state
  }
#pragma unroll
  for (int w = 0; w < W; ++w) y[i0 + w * 128] = h[w];
}
int main() {
  const int n = 148 * 5 * 128 * 2 * 8, iters = 64;
  double *x, *y; cudaMalloc(&x, 5ull * n * 8); cudaMalloc(&y, 1ull * n * 8);
  double* hx = new double[5ull * n];
  for (int i = 0; i < n; ++i) { hx[i] = -5.0 + (i % 97) * 0.1; hx[n + i] = 0.003 + (i % 13) * 1e-4; hx[2ull * n + i] = 88000.0 + i % 1000; hx[3ull * n + i] = 1.0 + (i % 7); hx[4ull * n + i] = 0.5; }
  cudaMemcpy(x, hx, 5ull * n * 8, cudaMemcpyHostToDevice);
  const size_t dyn = 40 * 1024;  // 5 blocks per SM
  cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
  cudaFuncSetAttribute(k<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int W = 1; W <= 2; ++W) for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(a);
    if (W == 1) k<1><<<n / 128, 128, dyn>>>(x, y, n, iters); else k<2><<<n / 256, 128, dyn>>>(x, y, n, iters);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    if (rep == 2) printf("W=%d: %.3f ms, %.2f G cell-steps/s (%s)\n", W, ms, (double)n * iters / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
