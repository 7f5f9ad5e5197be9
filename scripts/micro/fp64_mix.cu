// Microbenchmark: does a warp-level fp64 instruction (half rate: two cycles of the pipe) also hold the sub-partition's ISSUE
// port for two cycles?  Streams of 4 independent DFMA chains per thread with M independent FFMA / IMAD per DFMA interleaved.
// build here: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o hydrograd.jl_b200/fp64_mix_micro scripts/micro/fp64_mix.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int M, bool INT>
__global__ void k(double* out, int iters, double a, double b, float fa, float fb, int ia) {
  double x[4]; float y[4 * (M > 0 ? M : 1)]; int z[4 * (M > 0 ? M : 1)];
#pragma unroll
  for (int j = 0; j < 4; ++j) x[j] = threadIdx.x * 1e-3 + j;
#pragma unroll
  for (int j = 0; j < 4 * (M > 0 ? M : 1); ++j) { y[j] = threadIdx.x * 1e-3f + j; z[j] = threadIdx.x + j; }
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        x[j] = fma(x[j], a, b);
#pragma unroll
        for (int m = 0; m < M; ++m) {
          if (INT) z[j * M + m] = z[j * M + m] * ia + 7;
          else y[j * M + m] = fmaf(y[j * M + m], fa, fb);
        }
      }
  }
  const long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int j = 0; j < 4; ++j) s += x[j];
#pragma unroll
  for (int j = 0; j < 4 * (M > 0 ? M : 1); ++j) s += y[j] + z[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) reinterpret_cast<long long*>(out)[gridDim.x * blockDim.x] = t1 - t0;
}
template <int M, bool INT>
void run(double* d, int warps) {
  const int iters = 2000, threads = warps * 32;
  k<M, INT><<<148, threads>>>(d, iters, 1.0000001, 1e-9, 1.0001f, 1e-6f, 3);
  cudaDeviceSynchronize();
  long long cyc;
  cudaMemcpy(&cyc, reinterpret_cast<long long*>(d) + 148 * threads, 8, cudaMemcpyDeviceToHost);
  const double dfma = (double)iters * 16 * warps / 4;   // per sub-partition
  printf("warps/SM %2d  %d %s per DFMA : %.2f cycles per DFMA and sub-partition, %.2f instructions issued per cycle and sub-partition\n", warps, M,
         INT ? "IMAD" : "FFMA", cyc / dfma, dfma * (1 + M) / cyc);
}
int main() {
  double* d; cudaMalloc(&d, 148 * 1024 * 8 + 64);
  for (int w : {8, 16}) {
    run<0, false>(d, w); run<1, false>(d, w); run<2, false>(d, w); run<3, false>(d, w); run<1, true>(d, w); run<2, true>(d, w); run<3, true>(d, w);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
