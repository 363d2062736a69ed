// replay_kernels.cu — the off-policy trainers' replay buffer on the device (what SB3's ReplayBuffer.add / .sample do
// for train_td3.py:109-123), as two launches per call instead of a dozen PyTorch indexing kernels.
//
// A vector step of N envs stores N transitions; an update draws a batch of rows uniformly from what has been stored.
// Both are pure data movement, and in the TD3 loop both sit inside replayed CUDA graphs whose time is the latency of
// their chains of small dependent kernels (about 3 us per kernel): the write position, the number of rows stored and
// the sampler's stream position therefore live in DEVICE memory, so that identical launches serve every step.
//
//   sn_replay_add     rows (pos + r) % capacity <- (obs[r], next_obs[r], act[r], rew[r], done[r]);  optionally
//                     obs_carry[r] <- next_obs[r] (the loop's "current observation" buffer, which may BE `obs`);
//                     then pos <- (pos + n) % capacity, size <- min(size + n, capacity)   (second, one-thread launch)
//   sn_replay_sample  row b of the batch = stored row floor(u_b * size), u_b uniform in [0, 1) from the Philox stream
//                     (seed, *counter + b);  then *counter += batch                        (second, one-thread launch)
#include "../../include/supplynet.h"
#include "supplynet_error.h"
#include "philox.cuh"

#include <cuda_runtime.h>
#include <stdint.h>

namespace snrb {

constexpr int kBlock = 128;
constexpr int kRowsPerBlock = 4;          // one warp per transition

struct Rows {
  float* obs; float* next_obs; float* act; float* rew; float* done;
  int64_t capacity;
  int obs_dim, act_dim;
};

__global__ void __launch_bounds__(kBlock)
k_replay_add(const Rows rb, const int64_t* __restrict__ pos, int64_t n, const float* obs /* may alias obs_carry */,
             const float* __restrict__ next_obs, const float* __restrict__ act, const float* __restrict__ rew,
             const float* __restrict__ done, float* obs_carry) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * kRowsPerBlock + (threadIdx.x >> 5);
  if (r >= n) return;
  const int64_t dst = (*pos + r) % rb.capacity;
  const int od = rb.obs_dim, ad = rb.act_dim;
  // the same lane reads obs[r][c] and (when obs_carry aliases obs) overwrites it with next_obs[r][c]: read first
  for (int c = lane; c < od; c += 32) {
    const float o = obs[r * od + c], no = next_obs[r * od + c];
    rb.obs[dst * od + c] = o;
    rb.next_obs[dst * od + c] = no;
    if (obs_carry != nullptr) obs_carry[r * od + c] = no;
  }
  for (int c = lane; c < ad; c += 32) rb.act[dst * ad + c] = act[r * ad + c];
  if (lane == 0) {
    rb.rew[dst] = rew[r];
    rb.done[dst] = done != nullptr ? done[r] : 0.0f;
  }
}

__global__ void k_replay_advance(int64_t* pos, double* size, int64_t n, int64_t capacity) {
  *pos = (*pos + n) % capacity;
  const double s = *size + (double)n;
  *size = s < (double)capacity ? s : (double)capacity;
}

__global__ void __launch_bounds__(kBlock)
k_replay_sample(const Rows rb, const double* __restrict__ size, int64_t batch, uint64_t seed,
                const uint64_t* __restrict__ counter, float* __restrict__ obs, float* __restrict__ next_obs,
                float* __restrict__ act, float* __restrict__ rew, float* __restrict__ done) {
  const int lane = threadIdx.x & 31;
  const int64_t b = (int64_t)blockIdx.x * kRowsPerBlock + (threadIdx.x >> 5);
  if (b >= batch) return;
  // every lane of the warp computes the same draw (cheaper than a shuffle after one lane's ten rounds would be late)
  const uint64_t c = *counter + (uint64_t)b;
  const uint4 x = snr::philox4x32_10(make_uint4((uint32_t)c, (uint32_t)(c >> 32), 2u, 0u),
                                     make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  const double filled = *size;
  const double u = ((double)x.x * 4294967296.0 + (double)x.y) * (1.0 / 18446744073709551616.0);      // 64 bits in [0, 1)
  int64_t src = (int64_t)(u * filled);
  const int64_t last = (int64_t)filled - 1;
  src = src > last ? last : src;
  src = src < 0 ? 0 : src;
  const int od = rb.obs_dim, ad = rb.act_dim;
  for (int col = lane; col < od; col += 32) {
    obs[b * od + col] = rb.obs[src * od + col];
    next_obs[b * od + col] = rb.next_obs[src * od + col];
  }
  for (int col = lane; col < ad; col += 32) act[b * ad + col] = rb.act[src * ad + col];
  if (lane == 0) {
    rew[b] = rb.rew[src];
    done[b] = rb.done[src];
  }
}

__global__ void k_counter_bump(uint64_t* counter, uint64_t delta) { *counter += delta; }

inline int check(const sn_replay* rb, const char* who) {
  using sn_detail::fail;
  if (rb == nullptr) return fail(SN_ERR_INVALID, "%s: rb is NULL", who);
  if (!rb->obs || !rb->next_obs || !rb->act || !rb->rew || !rb->done || !rb->pos || !rb->size)
    return fail(SN_ERR_INVALID, "%s: every array of sn_replay must be non-NULL", who);
  if (rb->capacity < 1 || rb->obs_dim < 1 || rb->act_dim < 1)
    return fail(SN_ERR_INVALID, "%s: capacity=%lld obs_dim=%d act_dim=%d", who, (long long)rb->capacity, rb->obs_dim, rb->act_dim);
  return SN_OK;
}

inline Rows rows_of(const sn_replay* rb) {
  Rows r;
  r.obs = rb->obs; r.next_obs = rb->next_obs; r.act = rb->act; r.rew = rb->rew; r.done = rb->done;
  r.capacity = rb->capacity; r.obs_dim = rb->obs_dim; r.act_dim = rb->act_dim;
  return r;
}

}  // namespace snrb

extern "C" {

int32_t sn_replay_add(const sn_replay* rb, int64_t n, const float* obs, const float* next_obs, const float* act,
                      const float* rew, const float* done, float* obs_carry, void* stream) {
  using sn_detail::fail;
  int rc = snrb::check(rb, "sn_replay_add");
  if (rc != SN_OK) return rc;
  if (n < 1 || n > rb->capacity) return fail(SN_ERR_INVALID, "sn_replay_add: n=%lld with capacity=%lld", (long long)n, (long long)rb->capacity);
  if (!obs || !next_obs || !act || !rew) return fail(SN_ERR_INVALID, "sn_replay_add: obs, next_obs, act and rew must be non-NULL");
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned grid = (unsigned)((n + snrb::kRowsPerBlock - 1) / snrb::kRowsPerBlock);
  snrb::k_replay_add<<<grid, snrb::kBlock, 0, st>>>(snrb::rows_of(rb), rb->pos, n, obs, next_obs, act, rew, done, obs_carry);
  snrb::k_replay_advance<<<1, 1, 0, st>>>(rb->pos, rb->size, n, rb->capacity);
  return cudaGetLastError() == cudaSuccess ? SN_OK : fail(SN_ERR_CUDA, "sn_replay_add: kernel launch failed");
}

int32_t sn_replay_sample(const sn_replay* rb, int64_t batch, uint64_t seed, uint64_t* counter, float* obs, float* next_obs,
                         float* act, float* rew, float* done, void* stream) {
  using sn_detail::fail;
  int rc = snrb::check(rb, "sn_replay_sample");
  if (rc != SN_OK) return rc;
  if (batch < 1) return fail(SN_ERR_INVALID, "sn_replay_sample: batch=%lld", (long long)batch);
  if (!counter || !obs || !next_obs || !act || !rew || !done) return fail(SN_ERR_INVALID, "sn_replay_sample: counter and every output must be non-NULL");
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned grid = (unsigned)((batch + snrb::kRowsPerBlock - 1) / snrb::kRowsPerBlock);
  snrb::k_replay_sample<<<grid, snrb::kBlock, 0, st>>>(snrb::rows_of(rb), rb->size, batch, seed, counter, obs, next_obs, act, rew, done);
  snrb::k_counter_bump<<<1, 1, 0, st>>>(counter, (uint64_t)batch);
  return cudaGetLastError() == cudaSuccess ? SN_OK : fail(SN_ERR_CUDA, "sn_replay_sample: kernel launch failed");
}

}  // extern "C"
