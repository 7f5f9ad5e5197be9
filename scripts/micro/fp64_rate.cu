// Microbenchmark: what the fp64 pipe of one SM sustains (warp-level DFMA per cycle) as a function of resident warps and of the
// independent dependency chains per thread -- the quantity both tile kernels turn out to be bound by.
// build here: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o hydrograd.jl_b200/fp64_rate_micro scripts/micro/fp64_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void k(double* out, int iters, double a, double b) {
  double x[ILP];
#pragma unroll
  for (int j = 0; j < ILP; ++j) x[j] = threadIdx.x * 1e-3 + j;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int j = 0; j < ILP; ++j) x[j] = fma(x[j], a, b);
  }
  const long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int j = 0; j < ILP; ++j) s += x[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) reinterpret_cast<long long*>(out)[gridDim.x * blockDim.x] = t1 - t0;
}
template <int ILP>
void run(double* d, int warps) {
  const int iters = 2000, threads = warps * 32;   // one CTA per SM holds all the warps
  k<ILP><<<148, threads>>>(d, iters, 1.0000001, 1e-9);
  cudaDeviceSynchronize();
  long long cyc;
  cudaMemcpy(&cyc, reinterpret_cast<long long*>(d) + 148 * threads, 8, cudaMemcpyDeviceToHost);
  const double inst = (double)iters * 8 * ILP * warps;   // warp-level DFMA per SM
  printf("warps/SM %2d  chains/thread %d : %.3f warp-DFMA per cycle per SM (peak 2.0 = one per two cycles and sub-partition); %.1f cycles per DFMA per warp\n",
         warps, ILP, inst / cyc, (double)cyc / (iters * 8.0 * ILP));
}
int main() {
  double* d; cudaMalloc(&d, 148 * 1024 * 8 + 64);
  for (int w : {4, 8, 12, 16, 20, 32}) { run<1>(d, w); run<2>(d, w); run<4>(d, w); }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
