// gather.cu - score exchange of one ranking call that spans the GPUs of a box, without a collective.
//
// Reference call site: RecForYouProcess.java:56-59,92-94 (rank the candidates of one request,
// sort, cut) when the candidate list is sharded by rows over N GPUs: every
// rank needs all N score slices.  Instead of kernel -> NCCL all-gather, each rank's forward kernel
// stores its scores straight into its slice of EVERY rank's gather buffer (store_score(),
// common.cuh: peer memory mapped with CUDA IPC, stores travel over NVLink / NVSwitch), then one flag
// word per rank says "slice of step s is complete"; a consumer waits for N flags.
//
// One allocation per rank: [buffer 0: N x slice_rows floats][buffer 1: same][N flag words (256 B)].
// Two buffers alternate by step so that a rank one step ahead never overwrites scores a slower
// peer is still reading (a rank can only pass wait(s) once every peer has finished kernel s, i.e.
// has consumed step s - 1 in stream order).
#include <cstdio>
#include <cstring>
#include <new>

#include "kernels.h"

namespace srs {

struct PeerGather {
  int device = 0, world = 1, rank = 0;
  int64_t slice_rows = 0;
  size_t buf_bytes = 0;               // one buffer, 256-byte aligned
  uint8_t* local = nullptr;           // own allocation
  uint8_t* base[8] = {nullptr};       // every rank's allocation in this process' address space (own included)
  bool opened[8] = {false};
  unsigned int* counter = nullptr;    // local device words: [0] finished CTAs of the launch in flight, [1] steps signalled,
                                      // [2] steps waited for (device-side so that captured graphs can be replayed)
  uint32_t step = 0;                  // host copy: only its parity (which buffer) is used

  float* buffer(int r, int parity) const { return reinterpret_cast<float*>(base[r] + (size_t)parity * buf_bytes); }
  uint32_t* flags(int r) const { return reinterpret_cast<uint32_t*>(base[r] + 2 * buf_bytes); }
};

__global__ void gather_signal_kernel(unsigned int* counter, int n, uint32_t* f0, uint32_t* f1, uint32_t* f2,
                                     uint32_t* f3, uint32_t* f4, uint32_t* f5, uint32_t* f6, uint32_t* f7) {
  uint32_t* f[8] = {f0, f1, f2, f3, f4, f5, f6, f7};
  __shared__ uint32_t step_s;
  if (threadIdx.x == 0) step_s = ++counter[1];
  __syncthreads();
  __threadfence_system();
  if (threadIdx.x < n) *reinterpret_cast<volatile uint32_t*>(f[threadIdx.x]) = step_s;
}

// one warp: lane r polls the flag of rank r until it has reached this rank's own count of waits (wrap-safe compare)
__global__ void gather_wait_kernel(const uint32_t* flags, int n, unsigned int* counter) {
  __shared__ uint32_t step_s;
  if (threadIdx.x == 0) step_s = ++counter[2];
  __syncthreads();
  const uint32_t step = step_s;
  if (threadIdx.x < n) {
    const volatile uint32_t* f = flags + threadIdx.x;
    while ((int32_t)(*f - step) < 0) {
    }
  }
  __threadfence_system();
}

cudaError_t gather_create(int device, int world, int rank, int64_t slice_rows, PeerGather** out) {
  *out = nullptr;
  if (world < 1 || world > 8 || rank < 0 || rank >= world || slice_rows < 1) return cudaErrorInvalidValue;
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) return e;
  PeerGather* g = new (std::nothrow) PeerGather();
  if (!g) return cudaErrorMemoryAllocation;
  g->device = device; g->world = world; g->rank = rank; g->slice_rows = slice_rows;
  g->buf_bytes = (((size_t)world * slice_rows * 4) + 255) & ~(size_t)255;
  const size_t total = 2 * g->buf_bytes + 256;
  e = cudaMalloc(&g->local, total);
  if (e == cudaSuccess) e = cudaMemset(g->local, 0, total);
  if (e == cudaSuccess) e = cudaMalloc(&g->counter, 4 * sizeof(unsigned int));
  if (e == cudaSuccess) e = cudaMemset(g->counter, 0, 4 * sizeof(unsigned int));
  if (e != cudaSuccess) { cudaFree(g->local); cudaFree(g->counter); delete g; return e; }
  g->base[rank] = g->local;
  *out = g;
  return cudaSuccess;
}

cudaError_t gather_export(PeerGather* g, void* handle64) {
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaSetDevice(g->device);
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, g->local);
  if (e == cudaSuccess) memcpy(handle64, &h, 64);
  return e;
}

cudaError_t gather_connect(PeerGather* g, const void* handles) {
  cudaError_t e = cudaSetDevice(g->device);
  for (int r = 0; r < g->world && e == cudaSuccess; ++r) {
    if (r == g->rank || g->opened[r]) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, static_cast<const uint8_t*>(handles) + (size_t)r * 64, 64);
    void* p = nullptr;
    e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e == cudaSuccess) { g->base[r] = static_cast<uint8_t*>(p); g->opened[r] = true; }
  }
  return e;
}

void gather_destroy(PeerGather* g) {
  if (!g) return;
  cudaSetDevice(g->device);
  cudaDeviceSynchronize();
  for (int r = 0; r < g->world; ++r)
    if (g->opened[r]) cudaIpcCloseMemHandle(g->base[r]);
  cudaFree(g->local);
  cudaFree(g->counter);
  delete g;
}

bool gather_connected(const PeerGather* g) {
  for (int r = 0; r < g->world; ++r)
    if (!g->base[r]) return false;
  return true;
}

// Fill the exchange fields of a batch view for the next step; returns the step's buffer parity.
int gather_begin_step(PeerGather* g, BatchView& v, bool in_kernel_signal) {
  g->step += 1;
  const int parity = (int)(g->step & 1u);
  const size_t off = (size_t)g->rank * g->slice_rows;
  v.probs = g->buffer(g->rank, parity) + off;
  v.n_peers = 0;
  for (int r = 0; r < g->world; ++r)
    if (r != g->rank) v.peer_probs[v.n_peers++] = g->buffer(r, parity) + off;
  v.n_sig = 0;
  if (in_kernel_signal) {
    for (int r = 0; r < g->world; ++r) v.sig_flags[v.n_sig++] = g->flags(r) + g->rank;
    v.sig_counter = g->counter;
  }
  return parity;
}

cudaError_t gather_signal(PeerGather* g, cudaStream_t s) {
  uint32_t* f[8] = {nullptr};
  for (int r = 0; r < g->world; ++r) f[r] = g->flags(r) + g->rank;
  gather_signal_kernel<<<1, 32, 0, s>>>(g->counter, g->world, f[0], f[1], f[2], f[3], f[4], f[5], f[6], f[7]);
  ++g_launch_count;
  return cudaGetLastError();
}

cudaError_t gather_wait(PeerGather* g, cudaStream_t s) {
  gather_wait_kernel<<<1, 32, 0, s>>>(g->flags(g->rank), g->world, g->counter);
  ++g_launch_count;
  return cudaGetLastError();
}

float* gather_buffer(PeerGather* g, int parity) { return g->buffer(g->rank, parity & 1); }
int gather_parity(const PeerGather* g) { return (int)(g->step & 1u); }
int64_t gather_rows(const PeerGather* g) { return (int64_t)g->world * g->slice_rows; }
int gather_device(const PeerGather* g) { return g->device; }
int64_t gather_slice_rows(const PeerGather* g) { return g->slice_rows; }

}  // namespace srs
