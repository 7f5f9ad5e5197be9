// Library-owned halo exchange over NVLink peer memory: no NCCL, no Python in the data path.
//
// Every rank's context owns ONE cudaMalloc'd block: two parity copies of its receive buffer (layout of
// include/hydrograd_b200.h: per neighbour k a block of 6*n_k doubles [xi | q_x | q_y | l0 | l1 | l2]) followed by one
// 128-byte flag line per neighbour.  Peers map that block -- CUDA IPC across processes (one process per GPU, the
// torchrun / MPI case), plain peer access inside one process (one Julia process driving several GPUs) -- and an
// exchange is ONE small kernel per rank, k_comm_push: CTA k gathers the rank's cut cells of neighbour k and stores
// them straight into the peer's receive buffer through NVLink, then publishes the exchange's epoch in the peer's flag
// (st.release.sys).  The consumer is the fused RHS / VJP kernel itself: tiles are launched in band order (tiles without
// halo faces first), and only the band tiles spin on the flags (ld.acquire.sys) -- so the transfer overlaps the interior
// tiles inside a single launch and costs one ~3 us launch instead of pack + NCCL send/recv kernels + a second launch.
//
// Protocol (why two parity buffers and a monotone epoch are enough).  Rank A at epoch e: push_e writes B.recv[e&1], then
// B.flag[A] = e; A's consumer kernel_e waits for A.flag[B] >= e and reads A.recv[e&1].  push_{e+1} is stream-ordered after
// kernel_e, which saw flag >= e, i.e. B's push_e ran, which is stream-ordered after B's kernel_{e-1}: nobody still reads
// the buffer push_{e+1} overwrites.  A rank can be at most one epoch ahead of a neighbour.  An exchange is COLLECTIVE:
// every rank must issue the same sequence of exchanges.  A consumer that waits longer than the time-out (a peer that
// never pushed) raises the device error flag HG_ERR_COMM instead of hanging the GPU.
#include <sys/mman.h>
#include <sys/stat.h>
#include <fcntl.h>
#include <unistd.h>

#include <chrono>
#include <cstring>
#include <thread>

#include "hg_device.cuh"

namespace hg {
namespace {

constexpr int kFlagStride = 16;   // u64 per flag line (128 bytes)
constexpr uint32_t kMagic = 0x48474331u;   // "HGC1"

struct CommBlob {                 // what hg_comm_export writes (HG_COMM_HANDLE_BYTES = 128)
  cudaIpcMemHandle_t ipc;         // 64 bytes
  uint64_t self_ptr;              // the block's address in the exporting process (same-process peers use it directly)
  int64_t stride;                 // doubles per parity buffer
  int64_t flags_off;              // byte offset of the flag lines
  int32_t device, pid, n_nb;
  uint32_t magic;
  int64_t pad[3];
};
static_assert(sizeof(CommBlob) == HG_COMM_HANDLE_BYTES, "blob layout");

// CTA k: my cut cells towards neighbour k -> that peer's receive buffer, then the epoch into my flag line over there
__global__ void __launch_bounds__(512) k_comm_push(int32_t e0, int64_t Ns, const int32_t* __restrict__ bc_cell,
                                                   const int32_t* __restrict__ ptr, const double* __restrict__ Q,
                                                   const double* __restrict__ lam, double* const* __restrict__ dst,
                                                   unsigned long long* const* __restrict__ flag, unsigned long long epoch) {
  dev::pdl_launch_dependents();   // the consumer kernel (a programmatic dependent) may start: it needs the PEERS' pushes, not this one
  const int k = blockIdx.x;
  const int32_t p0 = ptr[k], n = ptr[k + 1] - p0;
  double* out = dst[k];
  for (int32_t i = threadIdx.x; i < n; i += blockDim.x) {
    const int32_t c = bc_cell[e0 + p0 + i];
    out[i] = Q[c]; out[n + i] = Q[Ns + c]; out[2 * n + i] = Q[2 * Ns + c];
    if (lam) { out[3 * n + i] = lam[c]; out[4 * n + i] = lam[Ns + c]; out[5 * n + i] = lam[2 * Ns + c]; }
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(flag[k]), "l"(epoch) : "memory");
}

int fail(hg_ctx* ctx, int code, const std::string& m) { ctx->err = m; return code; }

}  // namespace

// launched by fused_rhs / fused_vjp on comm-ready contexts (auto mode) and by hg_comm_exchange
int comm_push(hg_ctx* ctx, const double* d_Q, const double* d_lam) {
  hg_comm* cm = ctx->comm;
  if (!cm || !cm->connected) return fail(ctx, HG_ERR_STATE, "halo exchange: hg_comm_connect has not been called");
  cm->epoch++;
  const int par = (int)(cm->epoch & 1);
  k_comm_push<<<(unsigned)cm->n, 512, 0, ctx->stream>>>((int32_t)ctx->halo_e0, ctx->fh.Ns, ctx->fd.bc_cell.p, cm->d_ptr.p, d_Q, d_lam,
                                                        par ? cm->d_dst1.p : cm->d_dst0.p, cm->d_flag.p, cm->epoch);
  ctx->launches++;
  cm->pushed = true;
  return cudaGetLastError() == cudaSuccess ? HG_OK : fail(ctx, HG_ERR_CUDA, "k_comm_push launch failed");
}

}  // namespace hg

using hg::CommBlob;

static int comm_alloc(hg_ctx* ctx) {
  if (ctx->comm) return HG_OK;
  if (ctx->n_halo <= 0) { ctx->err = "hg_comm: the context has no halo boundaries"; return HG_ERR_ARG; }
  if (cudaSetDevice(ctx->opt.device) != cudaSuccess) { ctx->err = "cudaSetDevice"; return HG_ERR_CUDA; }
  hg_comm* cm = new hg_comm();
  cm->n = (int32_t)ctx->n_halo;
  cm->stride = 6 * ctx->n_halo_entries;
  cm->flags_off = ((2 * cm->stride * 8 + 127) / 128) * 128;
  cm->bytes = (size_t)cm->flags_off + (size_t)cm->n * hg::kFlagStride * 8;
  if (cudaMalloc(&cm->buf, cm->bytes) != cudaSuccess || cudaMemset(cm->buf, 0, cm->bytes) != cudaSuccess ||
      cudaDeviceSynchronize() != cudaSuccess) {
    delete cm;
    ctx->err = "hg_comm: cudaMalloc of the receive block failed";
    return HG_ERR_CUDA;
  }
  cm->recv[0] = (double*)cm->buf; cm->recv[1] = cm->recv[0] + cm->stride;
  cm->flags = (unsigned long long*)((char*)cm->buf + cm->flags_off);
  std::vector<int32_t> ptr(cm->n + 1, 0);
  for (int k = 0; k < cm->n; ++k) ptr[k + 1] = ptr[k] + (int32_t)ctx->bch.halo_counts[k];
  if (cm->d_ptr.upload(ptr, ctx->stream) != cudaSuccess) { cudaFree(cm->buf); delete cm; ctx->err = "hg_comm: upload"; return HG_ERR_CUDA; }
  ctx->comm = cm;
  return HG_OK;
}

extern "C" {

int hg_comm_export(hg_ctx* ctx, void* handle) {
  if (!ctx || !handle) return HG_ERR_ARG;
  if (ctx->opt.path == 1) { ctx->err = "halo exchange needs the fused path"; return HG_ERR_ARG; }
  int rc = comm_alloc(ctx);
  if (rc != HG_OK) return rc;
  hg_comm* cm = ctx->comm;
  CommBlob b;
  std::memset(&b, 0, sizeof(b));
  if (cudaIpcGetMemHandle(&b.ipc, cm->buf) != cudaSuccess) {
    cudaGetLastError();            // IPC can be unavailable (containers): same-process peers still work through self_ptr
    std::memset(&b.ipc, 0, sizeof(b.ipc));
  }
  b.self_ptr = (uint64_t)(uintptr_t)cm->buf; b.stride = cm->stride; b.flags_off = cm->flags_off;
  b.device = ctx->opt.device; b.pid = (int32_t)getpid(); b.n_nb = cm->n; b.magic = hg::kMagic;
  std::memcpy(handle, &b, sizeof(b));
  return HG_OK;
}

int hg_comm_connect(hg_ctx* ctx, int64_t n_neighbors, const void* peer_handles, const int64_t* peer_entry_offset,
                    const int64_t* peer_flag_index) {
  if (!ctx || !peer_handles || !peer_entry_offset || !peer_flag_index) return HG_ERR_ARG;
  int rc = comm_alloc(ctx);
  if (rc != HG_OK) return rc;
  hg_comm* cm = ctx->comm;
  if (cm->connected) { ctx->err = "hg_comm_connect: already connected (hg_comm_disconnect first)"; return HG_ERR_STATE; }
  if (n_neighbors != cm->n) { ctx->err = "hg_comm_connect: the context has " + std::to_string(cm->n) + " halo boundaries"; return HG_ERR_ARG; }
  if (cudaSetDevice(ctx->opt.device) != cudaSuccess) { ctx->err = "cudaSetDevice"; return HG_ERR_CUDA; }
  std::vector<double*> dst0(cm->n), dst1(cm->n);
  std::vector<unsigned long long*> flg(cm->n);
  for (int k = 0; k < cm->n; ++k) {
    CommBlob b;
    std::memcpy(&b, (const char*)peer_handles + (size_t)k * HG_COMM_HANDLE_BYTES, sizeof(b));
    if (b.magic != hg::kMagic) { ctx->err = "hg_comm_connect: handle " + std::to_string(k) + " was not written by hg_comm_export"; return HG_ERR_ARG; }
    if (peer_flag_index[k] < 0 || peer_flag_index[k] >= b.n_nb || peer_entry_offset[k] < 0 ||
        6 * (peer_entry_offset[k] + ctx->bch.halo_counts[k]) > b.stride) {
      ctx->err = "hg_comm_connect: block of neighbour " + std::to_string(k) + " does not fit the peer's receive buffer";
      return HG_ERR_ARG;
    }
    char* base = nullptr;
    if (b.pid == (int32_t)getpid()) {
      base = (char*)(uintptr_t)b.self_ptr;
      if (b.device != ctx->opt.device) {
        int can = 0;
        cudaDeviceCanAccessPeer(&can, ctx->opt.device, b.device);
        if (!can) { ctx->err = "hg_comm_connect: no peer access between devices " + std::to_string(ctx->opt.device) + " and " + std::to_string(b.device); return HG_ERR_CUDA; }
        const cudaError_t e = cudaDeviceEnablePeerAccess(b.device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { ctx->err = std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e); return HG_ERR_CUDA; }
        cudaGetLastError();
      }
    } else {
      void* p = nullptr;
      const cudaError_t e = cudaIpcOpenMemHandle(&p, b.ipc, cudaIpcMemLazyEnablePeerAccess);
      if (e != cudaSuccess) { cudaGetLastError(); ctx->err = std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(e); return HG_ERR_CUDA; }
      cm->opened.push_back(p);
      base = (char*)p;
    }
    dst0[k] = (double*)base + 6 * peer_entry_offset[k];
    dst1[k] = dst0[k] + b.stride;
    flg[k] = (unsigned long long*)(base + b.flags_off) + hg::kFlagStride * peer_flag_index[k];
  }
  if (cm->d_dst0.upload(dst0, ctx->stream) != cudaSuccess || cm->d_dst1.upload(dst1, ctx->stream) != cudaSuccess ||
      cm->d_flag.upload(flg, ctx->stream) != cudaSuccess) { ctx->err = "hg_comm_connect: upload"; return HG_ERR_CUDA; }
  cm->connected = true;
  return HG_OK;
}

int hg_comm_set_auto(hg_ctx* ctx, int32_t on) {
  if (!ctx || !ctx->comm) return HG_ERR_ARG;
  ctx->comm->auto_exchange = on != 0;
  return HG_OK;
}

int hg_comm_set_allreduce(hg_ctx* ctx, hg_allreduce_fn fn, void* user) {
  if (!ctx) return HG_ERR_ARG;
  ctx->allreduce = fn; ctx->allreduce_user = user;
  return HG_OK;
}

int hg_comm_exchange(hg_ctx* ctx, int32_t with_lambda) {
  if (!ctx) return HG_ERR_ARG;
  if (!ctx->state_set || (with_lambda && !ctx->lam_set)) { ctx->err = "hg_comm_exchange: state or lambda not set"; return HG_ERR_STATE; }
  if (cudaSetDevice(ctx->opt.device) != cudaSuccess) { ctx->err = "cudaSetDevice"; return HG_ERR_CUDA; }
  return hg::comm_push(ctx, ctx->fd.Q.p, with_lambda ? ctx->fd.lam.p : nullptr);
}

int hg_comm_disconnect(hg_ctx* ctx) {
  if (!ctx) return HG_ERR_ARG;
  hg_comm* cm = ctx->comm;
  if (!cm) return HG_OK;
  cudaSetDevice(ctx->opt.device);
  cudaStreamSynchronize(ctx->stream);
  for (void* p : cm->opened) cudaIpcCloseMemHandle(p);
  if (cm->buf) cudaFree(cm->buf);
  delete cm;
  ctx->comm = nullptr;
  return HG_OK;
}

// Rendezvous for hosts without their own messaging layer (a Julia / C driver with one process per GPU): a POSIX
// shared-memory segment named after the job carries every rank's handle and neighbour table.
int hg_comm_init_shm(hg_ctx* ctx, const char* job_name, int32_t rank, int32_t world, const int32_t* neighbor_ranks) {
  if (!ctx || !job_name || !neighbor_ranks || rank < 0 || rank >= world) return HG_ERR_ARG;
  constexpr int kMaxNb = 32;
  struct Slot { CommBlob blob; int32_t n_nb, pad; int32_t nb_rank[kMaxNb]; int64_t count[kMaxNb]; };
  struct Header { int arrived, connected; };
  int rc = comm_alloc(ctx);
  if (rc != HG_OK) return rc;
  hg_comm* cm = ctx->comm;
  if (cm->n > kMaxNb) { ctx->err = "hg_comm_init_shm: more than 32 neighbours"; return HG_ERR_ARG; }
  const std::string name = std::string("/hg_b200_") + job_name;
  const size_t bytes = sizeof(Header) + (size_t)world * sizeof(Slot);
  const int fd = shm_open(name.c_str(), O_CREAT | O_RDWR, 0600);
  if (fd < 0 || ftruncate(fd, (off_t)bytes) != 0) { if (fd >= 0) close(fd); ctx->err = "hg_comm_init_shm: shm_open failed"; return HG_ERR_ARG; }
  void* mem = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
  close(fd);
  if (mem == MAP_FAILED) { ctx->err = "hg_comm_init_shm: mmap failed"; return HG_ERR_ARG; }
  Header* hd = (Header*)mem;
  Slot* slots = (Slot*)((char*)mem + sizeof(Header));
  Slot& me = slots[rank];
  rc = hg_comm_export(ctx, &me.blob);
  auto wait_for = [&](int* counter, int target) {
    const auto t0 = std::chrono::steady_clock::now();
    while (__atomic_load_n(counter, __ATOMIC_ACQUIRE) < target) {
      if (std::chrono::steady_clock::now() - t0 > std::chrono::seconds(120)) return false;
      std::this_thread::sleep_for(std::chrono::milliseconds(1));
    }
    return true;
  };
  if (rc == HG_OK) {
    me.n_nb = cm->n;
    for (int k = 0; k < cm->n; ++k) { me.nb_rank[k] = neighbor_ranks[k]; me.count[k] = ctx->bch.halo_counts[k]; }
    __atomic_add_fetch(&hd->arrived, 1, __ATOMIC_ACQ_REL);
    if (!wait_for(&hd->arrived, world)) { rc = HG_ERR_STATE; ctx->err = "hg_comm_init_shm: timed out waiting for the other ranks"; }
  }
  if (rc == HG_OK) {
    std::vector<CommBlob> blobs(cm->n);
    std::vector<int64_t> off(cm->n), idx(cm->n);
    for (int k = 0; k < cm->n && rc == HG_OK; ++k) {
      const int32_t q = neighbor_ranks[k];
      if (q < 0 || q >= world || q == rank) { rc = HG_ERR_ARG; ctx->err = "hg_comm_init_shm: bad neighbour rank"; break; }
      const Slot& s = slots[q];
      int j = -1;
      int64_t before = 0;
      for (int m = 0; m < s.n_nb; ++m) { if (s.nb_rank[m] == rank) { j = m; break; } before += s.count[m]; }
      if (j < 0 || s.count[j] != ctx->bch.halo_counts[k]) {
        rc = HG_ERR_ARG;
        ctx->err = "hg_comm_init_shm: rank " + std::to_string(q) + " does not list this rank with the same cut";
        break;
      }
      blobs[k] = s.blob; off[k] = before; idx[k] = j;
    }
    if (rc == HG_OK) rc = hg_comm_connect(ctx, cm->n, blobs.data(), off.data(), idx.data());
  }
  // everybody has read the table (or failed) before the segment goes away
  const int done = __atomic_add_fetch(&hd->connected, 1, __ATOMIC_ACQ_REL);
  if (done == world) shm_unlink(name.c_str());
  else if (rc == HG_OK) wait_for(&hd->connected, world);
  munmap(mem, bytes);
  return rc;
}

}  // extern "C"
