// vecnorm_kernels.cu — device-side VecNormalize.
//
// The reference wraps its VecEnv in stable_baselines3.common.vec_env.VecNormalize(clip_obs=100,
// clip_reward=1000) (train_td3.py:100-101) and feeds HGE the un-normalised observation
// (hge.py:125).  SB3's source is not vendored in the reference and not installed here, so these
// kernels follow its documented algorithm — parity UNPINNED by reference tests:
//   RunningMeanStd(epsilon=1e-4): mean 0, var 1, count eps; update(x) merges the batch moments with
//     the parallel-variance formula (delta, M2);
//   step: obs_rms.update(obs); obs <- clip((obs-mean)/sqrt(var+eps), +-clip_obs);
//         ret <- ret*gamma + r; ret_rms.update(ret); r <- clip(r/sqrt(var_ret+eps), +-clip_reward);
//         ret[done] <- 0.
// Three launches: column moments (coalesced: a thread always sees the same column because the block
// size is a multiple of obs_dim), a one-block finalize that merges them into the running statistics,
// and an elementwise apply.
#include "../../include/supplynet.h"
#include "supplynet_error.h"

#include <cuda_runtime.h>
#include <cstdint>

namespace snv {

constexpr int kMaxOd = 56;      // 7 numbers x up to 8 facilities

// scratch layout: [0..od) sum x, [od..2od) sum x^2, [2od] sum ret, [2od+1] sum ret^2
__global__ void __launch_bounds__(256)
k_moments(int64_t n, int od, int tpb, const float* __restrict__ obs, const float* __restrict__ reward, double gamma,
          double* __restrict__ returns, double* __restrict__ scratch) {
  __shared__ double sh[2 * 256];
  const int tid = threadIdx.x;
  double sx = 0.0, sxx = 0.0;
  if (tid < tpb) {
    // block b covers rows [b*rows_per_block, ...), thread = (row_in_tile, column)
    const int64_t total = n * od;
    for (int64_t i = (int64_t)blockIdx.x * tpb + tid; i < total; i += (int64_t)gridDim.x * tpb) {
      const double x = (double)__ldg(obs + i);
      sx += x;
      sxx += x * x;
    }
  }
  sh[tid] = sx;
  sh[256 + tid] = sxx;
  __syncthreads();
  if (tid < od) {
    double a = 0.0, b = 0.0;
    for (int t = tid; t < tpb; t += od) { a += sh[t]; b += sh[256 + t]; }
    atomicAdd(&scratch[tid], a);
    atomicAdd(&scratch[od + tid], b);
  }
  if (reward != nullptr) {
    double sr = 0.0, srr = 0.0;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + tid; e < n; e += (int64_t)gridDim.x * blockDim.x) {
      const double r = returns[e] * gamma + (double)__ldg(reward + e);
      returns[e] = r;
      sr += r;
      srr += r * r;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      sr += __shfl_xor_sync(0xffffffffu, sr, o);
      srr += __shfl_xor_sync(0xffffffffu, srr, o);
    }
    __syncthreads();
    if ((tid & 31) == 0) { sh[tid >> 5] = sr; sh[256 + (tid >> 5)] = srr; }
    __syncthreads();
    if (tid == 0) {
      double a = 0.0, b = 0.0;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { a += sh[w]; b += sh[256 + w]; }
      atomicAdd(&scratch[2 * od], a);
      atomicAdd(&scratch[2 * od + 1], b);
    }
  }
}

__device__ __forceinline__ void merge(double mean, double var, double count, double bmean, double bvar, double bcount,
                                      double& nmean, double& nvar, double& ncount) {
  const double delta = bmean - mean;
  const double tot = count + bcount;
  nmean = mean + delta * bcount / tot;
  const double m2 = var * count + bvar * bcount + delta * delta * count * bcount / tot;
  nvar = m2 / tot;
  ncount = tot;
}

// rms layout: obs_rms = mean[od], var[od], count ; ret_rms = mean, var, count.  In and out buffers may be the
// same (every input is read before the barrier, every output written after it).
__global__ void k_finalize(int64_t n, int od, int has_reward, const double* scratch,
                           const double* obs_in, double* obs_out, const double* ret_in, double* ret_out) {
  const int c = threadIdx.x;
  const double bn = (double)n;
  double m = 0.0, v = 0.0, k = 0.0;
  if (c < od) {
    const double bmean = scratch[c] / bn;
    const double bvar = fmax(scratch[od + c] / bn - bmean * bmean, 0.0);
    merge(obs_in[c], obs_in[od + c], obs_in[2 * od], bmean, bvar, bn, m, v, k);
  }
  double rm = 0.0, rv = 0.0, rk = 0.0;
  if (c == 63) {                               // obs_dim <= 56: thread 63 never owns a column
    if (has_reward) {
      const double bmean = scratch[2 * od] / bn;
      const double bvar = fmax(scratch[2 * od + 1] / bn - bmean * bmean, 0.0);
      merge(ret_in[0], ret_in[1], ret_in[2], bmean, bvar, bn, rm, rv, rk);
    } else {
      rm = ret_in[0]; rv = ret_in[1]; rk = ret_in[2];
    }
  }
  __syncthreads();
  if (c < od) {
    obs_out[c] = m;
    obs_out[od + c] = v;
    if (c == 0) obs_out[2 * od] = k;
  }
  if (c == 63) { ret_out[0] = rm; ret_out[1] = rv; ret_out[2] = rk; }
}

__global__ void __launch_bounds__(256)
k_apply(int64_t n, int od, int tpb, const float* __restrict__ obs, float* __restrict__ obs_out,
        const float* __restrict__ reward, float* __restrict__ reward_out, const double* __restrict__ obs_rms,
        const double* __restrict__ ret_rms, float clip_obs, float clip_reward, double eps, int reset_returns,
        double* __restrict__ returns) {
  const int tid = threadIdx.x;
  if (obs != nullptr && tid < tpb) {
    const int c = tid % od;
    const float mean = (float)obs_rms[c];
    const float istd = (float)(1.0 / sqrt(obs_rms[od + c] + eps));
    const int64_t total = n * od;
    for (int64_t i = (int64_t)blockIdx.x * tpb + tid; i < total; i += (int64_t)gridDim.x * tpb) {
      const float z = (__ldg(obs + i) - mean) * istd;
      obs_out[i] = fminf(fmaxf(z, -clip_obs), clip_obs);
    }
  }
  if (reward != nullptr) {
    const float istd = (float)(1.0 / sqrt(ret_rms[1] + eps));
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + tid; e < n; e += (int64_t)gridDim.x * blockDim.x) {
      const float z = __ldg(reward + e) * istd;
      reward_out[e] = fminf(fmaxf(z, -clip_reward), clip_reward);
      if (reset_returns) returns[e] = 0.0;
    }
  }
}

}  // namespace snv

namespace {
int threads_for(int od) { return (256 / od) * od; }
unsigned blocks_for(int64_t n, int od, int tpb) {
  int64_t b = (n * od + (int64_t)tpb * 8 - 1) / ((int64_t)tpb * 8);
  if (b < 1) b = 1;
  if (b > 148 * 8) b = 148 * 8;      // a few waves of the 148 SMs; grid-stride beyond
  return (unsigned)b;
}
}  // namespace

extern "C" {

int32_t sn_vecnorm_update(int64_t n_envs, int32_t obs_dim, const float* obs, const float* reward, double gamma,
                          double* returns, const double* obs_rms_in, double* obs_rms_out, const double* ret_rms_in,
                          double* ret_rms_out, double* scratch, void* stream) {
  using sn_detail::fail;
  if (n_envs < 1 || obs_dim < 1 || obs_dim > snv::kMaxOd)
    return fail(SN_ERR_INVALID, "sn_vecnorm_update: n_envs=%lld, obs_dim=%d (1..%d)", (long long)n_envs, obs_dim, snv::kMaxOd);
  if (!obs || !obs_rms_in || !obs_rms_out || !ret_rms_in || !ret_rms_out || !scratch)
    return fail(SN_ERR_INVALID, "sn_vecnorm_update: obs, the four statistics buffers and scratch must be non-NULL");
  if (reward && !returns) return fail(SN_ERR_INVALID, "sn_vecnorm_update: returns must be given with reward");
  cudaStream_t st = (cudaStream_t)stream;
  if (cudaMemsetAsync(scratch, 0, sizeof(double) * (2 * obs_dim + 2), st) != cudaSuccess)
    return fail(SN_ERR_CUDA, "sn_vecnorm_update: clearing scratch failed");
  const int tpb = threads_for(obs_dim);
  snv::k_moments<<<blocks_for(n_envs, obs_dim, tpb), 256, 0, st>>>(n_envs, obs_dim, tpb, obs, reward, gamma, returns, scratch);
  snv::k_finalize<<<1, 64, 0, st>>>(n_envs, obs_dim, reward != nullptr, scratch, obs_rms_in, obs_rms_out, ret_rms_in, ret_rms_out);
  return cudaGetLastError() == cudaSuccess ? SN_OK : sn_detail::fail(SN_ERR_CUDA, "sn_vecnorm: kernel launch failed");
}

int32_t sn_vecnorm_apply(int64_t n_envs, int32_t obs_dim, const float* obs, float* obs_out, const float* reward,
                         float* reward_out, const double* obs_rms, const double* ret_rms, float clip_obs,
                         float clip_reward, double epsilon, int32_t reset_returns, double* returns, void* stream) {
  using sn_detail::fail;
  if (n_envs < 1 || obs_dim < 1 || obs_dim > snv::kMaxOd)
    return fail(SN_ERR_INVALID, "sn_vecnorm_apply: n_envs=%lld, obs_dim=%d (1..%d)", (long long)n_envs, obs_dim, snv::kMaxOd);
  if ((obs && (!obs_out || !obs_rms)) || (reward && (!reward_out || !ret_rms)))
    return fail(SN_ERR_INVALID, "sn_vecnorm_apply: an input was given without its output / statistics buffer");
  if (reward && reset_returns && !returns) return fail(SN_ERR_INVALID, "sn_vecnorm_apply: reset_returns needs returns");
  const int tpb = threads_for(obs_dim);
  snv::k_apply<<<blocks_for(n_envs, obs_dim, tpb), 256, 0, (cudaStream_t)stream>>>(
      n_envs, obs_dim, tpb, obs, obs_out, reward, reward_out, obs_rms, ret_rms, clip_obs, clip_reward, epsilon,
      reset_returns, returns);
  return cudaGetLastError() == cudaSuccess ? SN_OK : sn_detail::fail(SN_ERR_CUDA, "sn_vecnorm: kernel launch failed");
}

}  // extern "C"
