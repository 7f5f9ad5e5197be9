// Microbenchmark: what can the host link of this box carry in the shapes the host-buffer pipeline (hg_rhs / hg_rhs_vjp with
// host pointers) uses?  384 MB in + 384 MB out (the state and dQdt of the 16M-cell mesh), pinned memory, one GPU:
//   whole buffers, one direction at a time and both at once;
//   K chunks, each one strided copy of 3 rows (cudaMemcpy2DAsync) or 3 contiguous copies, H2D on one stream, the D2H of chunk c
//   on another stream released by an event after chunk c + lag has landed (what the tiles' readiness does in the library).
// build here (no GPU needed), run on the box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o hydrograd.jl_b200/pcie_pipeline_micro scripts/micro/pcie_pipeline.cu
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)
using clk = std::chrono::steady_clock;
static double ms_since(clk::time_point t0) { return std::chrono::duration<double, std::milli>(clk::now() - t0).count(); }

int main(int argc, char** argv) {
  const size_t N = argc > 1 ? (size_t)atoll(argv[1]) : 16000000;
  const size_t bytes = 3 * N * 8;
  double *hin, *hout, *din, *dout;
  CK(cudaMallocHost(&hin, bytes)); CK(cudaMallocHost(&hout, bytes));
  CK(cudaMalloc(&din, bytes)); CK(cudaMalloc(&dout, bytes));
  for (size_t i = 0; i < 3 * N; ++i) hin[i] = (double)i;
  CK(cudaMemset(dout, 0, bytes));
  cudaStream_t si, so;
  CK(cudaStreamCreateWithFlags(&si, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&so, cudaStreamNonBlocking));
  const double gb = bytes / 1e9;
  auto run = [&](const char* name, auto&& fn) {
    fn(); CK(cudaDeviceSynchronize());
    double best = 1e30;
    for (int r = 0; r < 5; ++r) {
      auto t0 = clk::now();
      fn(); CK(cudaDeviceSynchronize());
      const double ms = ms_since(t0);
      if (ms < best) best = ms;
    }
    printf("%-58s %7.2f ms  %6.1f GB/s per direction\n", name, best, gb / best * 1e3);
  };
  run("H2D whole buffer", [&] { CK(cudaMemcpyAsync(din, hin, bytes, cudaMemcpyHostToDevice, si)); });
  run("D2H whole buffer", [&] { CK(cudaMemcpyAsync(hout, dout, bytes, cudaMemcpyDeviceToHost, so)); });
  run("both whole buffers at once", [&] {
    CK(cudaMemcpyAsync(din, hin, bytes, cudaMemcpyHostToDevice, si));
    CK(cudaMemcpyAsync(hout, dout, bytes, cudaMemcpyDeviceToHost, so));
  });
  std::vector<cudaEvent_t> ev(256);
  for (auto& e : ev) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  for (int K : (N % 64 != 0 ? std::vector<int>{30} : std::vector<int>{8, 16, 32, 64, 128})) {
    for (int strided = 1; strided >= 0; --strided) {
      for (int lag : {1, 2}) {
        char name[128];
        snprintf(name, sizeof name, "K = %3d chunks, %s, D2H of chunk c after chunk c+%d", K, strided ? "one 2-D copy of 3 rows" : "3 contiguous copies   ", lag);
        run(name, [&] {
          const size_t csz = (N + K - 1) / K;
          auto copy = [&](double* dst, const double* src, size_t r0, size_t r1, cudaMemcpyKind kind, cudaStream_t s) {
            if (strided) CK(cudaMemcpy2DAsync(dst + r0, N * 8, src + r0, N * 8, (r1 - r0) * 8, 3, kind, s));
            else for (int q = 0; q < 3; ++q) CK(cudaMemcpyAsync(dst + q * N + r0, src + q * N + r0, (r1 - r0) * 8, kind, s));
          };
          for (int c = 0; c < K; ++c) {
            const size_t r0 = c * csz, r1 = r0 + csz < N ? r0 + csz : N;
            copy(din, hin, r0, r1, cudaMemcpyHostToDevice, si);
            CK(cudaEventRecord(ev[c], si));
          }
          for (int c = 0; c < K; ++c) {
            const size_t r0 = c * csz, r1 = r0 + csz < N ? r0 + csz : N;
            CK(cudaStreamWaitEvent(so, ev[c + lag < K ? c + lag : K - 1], 0));
            copy(hout, dout, r0, r1, cudaMemcpyDeviceToHost, so);
          }
        });
      }
    }
  }
  // ---- the library's real shape: N is not a multiple of anything, so the rows of a chunk start at 8-byte aligned addresses only.
  // align = 8: chunk boundaries at c * ceil(N / K) (what hg_rhs did); align > 8: three copies per chunk whose boundaries are shifted
  // per row so that every HOST address is a multiple of `align` bytes (the device side keeps the same cell offsets).
  if (N % 64 != 0) {
    for (int K : {30, 32}) {
      for (size_t align : {(size_t)8, (size_t)64, (size_t)256, (size_t)4096}) {
        char name[128];
        snprintf(name, sizeof name, "N = %zu, K = %d, 3 copies, host addresses %% %zu == 0", N, K, align);
        run(name, [&] {
          const size_t a = align / 8;                                  // cells
          const size_t csz = ((N + K - 1) / K + a - 1) / a * a;
          auto bounds = [&](const double* hrow, int c, size_t& r0, size_t& r1) {   // rows [r0, r1) of chunk c for the row starting at hrow
            const size_t sh = ((size_t)hrow / 8) % a;                    // hrow + 8 r is aligned when (sh + r) % a == 0
            const size_t lo = (size_t)c * csz, hi = lo + csz;
            r0 = c == 0 ? 0 : (lo >= sh ? lo - sh : 0);
            r1 = c == K - 1 ? N : (hi >= sh ? hi - sh : 0);
            if (r1 > N) r1 = N;
            if (r0 > r1) r0 = r1;
          };
          for (int c = 0; c < K; ++c) {
            for (int q = 0; q < 3; ++q) {
              size_t r0, r1; bounds(hin + q * N, c, r0, r1);
              if (r1 > r0) CK(cudaMemcpyAsync(din + q * N + r0, hin + q * N + r0, (r1 - r0) * 8, cudaMemcpyHostToDevice, si));
            }
            CK(cudaEventRecord(ev[c], si));
          }
          for (int c = 0; c < K; ++c) {
            CK(cudaStreamWaitEvent(so, ev[c + 1 < K ? c + 1 : K - 1], 0));
            for (int q = 0; q < 3; ++q) {
              size_t r0, r1; bounds(hout + q * N, c, r0, r1);
              if (r1 > r0) CK(cudaMemcpyAsync(hout + q * N + r0, dout + q * N + r0, (r1 - r0) * 8, cudaMemcpyDeviceToHost, so));
            }
          }
        });
      }
    }
  }
  return 0;
}
