// Microbenchmark: what does a thread pay to issue N cp.async.bulk copies of 2 KB, and when does the data land?
// build here (no GPU needed), run on the box:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o hydrograd.jl_b200/bulk_issue_micro scripts/micro/bulk_issue.cu
// (the executable travels with the snapshot; it is git-ignored)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int N, int ROWB>
__global__ void k(const char* src, size_t stride_cta, int iters, unsigned long long* out, int pad_smem) {
  extern __shared__ __align__(128) unsigned char sm[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(sm);
  char* dst = reinterpret_cast<char*>(sm) + 128;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(bar)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  unsigned long long t_issue = 0, t_wait = 0;
  uint32_t parity = 0;
  for (int it = 0; it < iters; ++it) {
    const char* s = src + (((size_t)blockIdx.x * iters + it) * stride_cta) % ((size_t)96 << 20);   // 96 MB footprint: L2-resident on the second pass
    if (threadIdx.x == 0) {
      const long long t0 = clock64();
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bar)), "r"(N * ROWB) : "memory");
#pragma unroll
      for (int j = 0; j < N; ++j)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(dst + j * ROWB)),
                     "l"(s + (size_t)j * ROWB), "r"(ROWB), "r"(s32(bar)) : "memory");
      const long long t1 = clock64();
      uint32_t done;
      do {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(s32(bar)), "r"(parity) : "memory");
      } while (!done);
      const long long t2 = clock64();
      t_issue += t1 - t0; t_wait += t2 - t1;
    }
    parity ^= 1;
    __syncthreads();
  }
  if (threadIdx.x == 0) { atomicAdd(out, t_issue); atomicAdd(out + 1, t_wait); }
}
template <int N, int ROWB>
void run(const char* d, unsigned long long* dout, int ctas_per_sm, const char* what) {
  const int iters = 64, grid = 148 * ctas_per_sm;
  const size_t stride = (size_t)N * ROWB;
  const int smem = 227 * 1024 / ctas_per_sm - 1024;   // forces the co-residency
  cudaFuncSetAttribute(k<N, ROWB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int rep = 0; rep < 2; ++rep) {   // second pass: sources (partly) resident in L2
    cudaMemset(dout, 0, 16);
    k<N, ROWB><<<grid, 128, smem>>>(d, stride, iters, dout, 0);
    cudaDeviceSynchronize();
    unsigned long long h[2];
    cudaMemcpy(h, dout, 16, cudaMemcpyDeviceToHost);
    const double den = (double)grid * iters;
    printf("%-10s N=%2d x %5d B  CTAs/SM=%d  %s: issue %.0f cycles (%.0f per copy), then wait %.0f\n", what, N, ROWB, ctas_per_sm, rep ? "L2  " : "HBM ",
           h[0] / den, h[0] / den / N, h[1] / den);
  }
}
int main() {
  char* d; unsigned long long* dout;
  cudaMalloc(&d, (size_t)97 << 20); cudaMemset(d, 1, (size_t)97 << 20); cudaMalloc(&dout, 16);
  for (int c : {1, 3, 5}) {
    run<1, 2048>(d, dout, c, "rows");  run<4, 2048>(d, dout, c, "rows"); run<8, 2048>(d, dout, c, "rows"); run<16, 2048>(d, dout, c, "rows");
    run<1, 32768>(d, dout, c, "one block"); run<4, 8192>(d, dout, c, "blocks");
  }
  cudaError_t e = cudaGetLastError();
  printf("%s\n", cudaGetErrorString(e));
  return 0;
}
