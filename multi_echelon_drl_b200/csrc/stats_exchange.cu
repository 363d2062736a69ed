// stats_exchange.cu — the path's only collective as ONE kernel over peer memory (NVLink / NVSwitch).
//
// Envs never interact, so ranks exchange nothing but the 16-double statistics vector the step / rollout kernels keep
// accumulating (all-reduce(sum) once per log interval).  Through NCCL that is a copy kernel, an event
// hand-over to NCCL's stream and a collective kernel that does not fit next to the rollout's blocks: 25-30 us per call
// around a 150 us episode kernel.  Here every rank owns a small MAILBOX in device memory that all peers map
// (cudaIpc*, peer access enabled lazily by the driver); one launch of 16 x world threads
//   1. stores its 16 values into slot [parity][rank] of EVERY rank's mailbox - plain stores that travel over NVLink,
//      each 16-byte entry carrying value and sequence tag together (the low-latency protocol of collective
//      libraries: 8-byte accesses are single-copy atomic, so a half whose tag matches carries that call's data -
//      no fence, no flag word, no second trip), and
//   2. polls its OWN mailbox (local memory) until the slots of all ranks carry this call's tag, then adds them up in
//      rank order: every rank computes bit-identical sums.
// Two slot parities suffice: a rank can only reach call k+2 after it has collected call k+1, which needs every
// peer's contribution to k+1, which that peer issued after it finished collecting call k (stream order).
#include "../../include/supplynet.h"
#include "supplynet_error.h"

#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

using sn_detail::fail;

namespace snx {

constexpr int kWords = SN_STATS_WORDS;

__device__ __forceinline__ void ll_store(double* entry, double v, uint32_t tag) {
  const unsigned long long bits = (unsigned long long)__double_as_longlong(v), t = (unsigned long long)tag << 32;
  const unsigned long long w0 = (bits & 0xffffffffull) | t, w1 = (bits >> 32) | t;
  asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(entry), "l"(w0), "l"(w1) : "memory");
}
__device__ __forceinline__ bool ll_poll(const double* entry, uint32_t tag, double& v) {
  unsigned long long w0, w1;
  asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(entry) : "memory");
  v = __longlong_as_double((long long)((w0 & 0xffffffffull) | (w1 << 32)));
  return (uint32_t)(w0 >> 32) == tag && (uint32_t)(w1 >> 32) == tag;
}

// thread = (rank p, word i)
__global__ void __launch_bounds__(kWords * SN_MAIL_MAX_RANKS)
k_stats_allreduce(const double* in, double* out, void* const* __restrict__ mails, int rank, int world, uint32_t seq) {
  __shared__ double part[SN_MAIL_MAX_RANKS][kWords];
  const int i = threadIdx.x % kWords, p = threadIdx.x / kWords;
  const size_t par = (size_t)(seq & 1u) * world;
  if (p < world) {
    ll_store(reinterpret_cast<double*>(mails[p]) + ((par + rank) * kWords + i) * 2, in[i], seq);
    const double* src = reinterpret_cast<const double*>(mails[rank]) + ((par + p) * kWords + i) * 2;
    double v;
    const long long t0 = clock64();
    while (!ll_poll(src, seq, v))
      if (clock64() - t0 > (60ll << 30)) __trap();          // about half a minute: a peer that never arrives is an error, not a hang
    part[p][i] = v;
  }
  __syncthreads();
  if (threadIdx.x < kWords) {
    double s = 0.0;
    for (int q = 0; q < world; ++q) s += part[q][threadIdx.x];      // rank order: identical on every rank
    out[threadIdx.x] = s;
  }
}

static int cuda_fail(const char* what, cudaError_t e) { return fail(SN_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e)); }

}  // namespace snx

extern "C" {

int64_t sn_mail_bytes(int32_t world) {
  if (world < 1 || world > SN_MAIL_MAX_RANKS) return fail(SN_ERR_INVALID, "sn_mail_bytes: world=%d outside 1..%d", world, SN_MAIL_MAX_RANKS);
  return (int64_t)2 * world * snx::kWords * 16;
}

int32_t sn_mail_alloc(int32_t world, void** mail, unsigned char* handle) {
  if (mail == nullptr || handle == nullptr) return fail(SN_ERR_INVALID, "sn_mail_alloc: mail and handle must be non-NULL");
  const int64_t bytes = sn_mail_bytes(world);
  if (bytes < 0) return (int32_t)bytes;
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, (size_t)bytes);
  if (e != cudaSuccess) return snx::cuda_fail("sn_mail_alloc cudaMalloc", e);
  if ((e = cudaMemset(p, 0, (size_t)bytes)) != cudaSuccess || (e = cudaDeviceSynchronize()) != cudaSuccess) {
    cudaFree(p);
    return snx::cuda_fail("sn_mail_alloc cudaMemset", e);
  }
  cudaIpcMemHandle_t h;
  static_assert(sizeof(h) == SN_MAIL_HANDLE_BYTES, "handle size");
  if ((e = cudaIpcGetMemHandle(&h, p)) != cudaSuccess) {
    cudaGetLastError();
    memset(handle, 0, SN_MAIL_HANDLE_BYTES);      // usable inside this process (sn_enable_peer_access), not across processes
  } else {
    memcpy(handle, &h, SN_MAIL_HANDLE_BYTES);
  }
  *mail = p;
  return SN_OK;
}

int32_t sn_mail_open(const unsigned char* handle, void** peer_mail) {
  if (handle == nullptr || peer_mail == nullptr) return fail(SN_ERR_INVALID, "sn_mail_open: handle and peer_mail must be non-NULL");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, SN_MAIL_HANDLE_BYTES);
  void* p = nullptr;
  const cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) { cudaGetLastError(); return snx::cuda_fail("sn_mail_open cudaIpcOpenMemHandle", e); }
  *peer_mail = p;
  return SN_OK;
}

int32_t sn_mail_close(void* peer_mail) {
  const cudaError_t e = cudaIpcCloseMemHandle(peer_mail);
  if (e != cudaSuccess) { cudaGetLastError(); return snx::cuda_fail("sn_mail_close", e); }
  return SN_OK;
}

int32_t sn_mail_free(void* mail) {
  const cudaError_t e = cudaFree(mail);
  if (e != cudaSuccess) { cudaGetLastError(); return snx::cuda_fail("sn_mail_free", e); }
  return SN_OK;
}

int32_t sn_enable_peer_access(int32_t device, int32_t peer) {
  int cur = 0, can = 0;
  cudaError_t e = cudaGetDevice(&cur);
  if (e != cudaSuccess) return snx::cuda_fail("sn_enable_peer_access", e);
  if (device == peer) return SN_OK;
  if ((e = cudaDeviceCanAccessPeer(&can, device, peer)) != cudaSuccess) { cudaGetLastError(); return snx::cuda_fail("cudaDeviceCanAccessPeer", e); }
  if (!can) return fail(SN_ERR_UNSUPPORTED, "sn_enable_peer_access: device %d cannot map memory of device %d", device, peer);
  cudaSetDevice(device);
  e = cudaDeviceEnablePeerAccess(peer, 0);
  cudaSetDevice(cur);
  if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); return snx::cuda_fail("cudaDeviceEnablePeerAccess", e); }
  cudaGetLastError();
  return SN_OK;
}

int32_t sn_stats_allreduce(const double* stats_in, double* stats_out, void* const* mails, int32_t rank, int32_t world,
                           uint32_t seq, void* stream) {
  if (stats_in == nullptr || stats_out == nullptr || mails == nullptr) return fail(SN_ERR_INVALID, "sn_stats_allreduce: NULL argument");
  if (world < 1 || world > SN_MAIL_MAX_RANKS || rank < 0 || rank >= world)
    return fail(SN_ERR_INVALID, "sn_stats_allreduce: rank %d / world %d (at most %d ranks)", rank, world, SN_MAIL_MAX_RANKS);
  if (seq == 0) return fail(SN_ERR_INVALID, "sn_stats_allreduce: seq counts from 1 (0 is the tag of an empty mailbox)");
  snx::k_stats_allreduce<<<1, snx::kWords * world, 0, (cudaStream_t)stream>>>(stats_in, stats_out, mails, rank, world, seq);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return snx::cuda_fail("sn_stats_allreduce launch", e);
  return SN_OK;
}

}  // extern "C"
