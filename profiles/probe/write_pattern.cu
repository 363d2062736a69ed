// Synthetic probe: how fast can 65,536 threads (one per env, like k_rollout) stream a [T][N][28] float
// trajectory into HBM, with nothing else going on?  Variants:
//   0  time-major slabs (the trajectory layout): at period s every warp writes its 3,584-byte chunk of slab s
//   1  warp-major (each warp's 100 chunks contiguous in memory): same bytes, different addresses
//   2  time-major, but 4x the threads (each thread writes 28 B... i.e. quarter chunks) to show the effect of
//      more warps in flight
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -shared -Xcompiler -fPIC -o libwrite_pattern.so write_pattern.cu
#include <cuda_runtime.h>
#include <stdint.h>

// variant 3/4: like 0 plus the kernel's input stream (16-byte action row + 4-byte demand per env and period),
// prefetched one period (3) or `PD` periods (4, register ring, fully unrolled) ahead and folded into the chain
constexpr int PD = 4;
__global__ void __launch_bounds__(64) k_slab_in(float* __restrict__ out, const int4* __restrict__ act,
                                                 const int* __restrict__ dem, int64_t n, int T, int variant, int delay) {
  const int64_t env = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (env >= n) return;
  const int lane = threadIdx.x & 31;
  const int64_t warp_env0 = env - lane;
  float acc = (float)env;
  if (variant == 3) {
    int4 a_pre = __ldg(act + env);
    int d_pre = __ldg(dem + n + env);
    for (int s = 0; s < T; ++s) {
      const int4 a = a_pre; const int d = d_pre;
      if (s + 1 < T) { a_pre = __ldg(act + (int64_t)(s + 1) * n + env); d_pre = __ldg(dem + (int64_t)(s + 2) * n + env); }
      acc += (float)(a.x + a.y + a.z + a.w + d);
      for (int i = 0; i < delay; ++i) acc = acc * 1.0001f + 0.5f;
      float4 v = make_float4(acc, acc, acc, acc);
      float4* base = reinterpret_cast<float4*>(out + ((int64_t)s * n + warp_env0) * 28);
#pragma unroll
      for (int j = 0; j < 7; ++j) base[j * 32 + lane] = v;
    }
  } else {
    int4 a_pre[PD]; int d_pre[PD];
#pragma unroll
    for (int j = 0; j < PD; ++j) { a_pre[j] = __ldg(act + (int64_t)j * n + env); d_pre[j] = __ldg(dem + (int64_t)(j + 1) * n + env); }
    for (int s0 = 0; s0 < T; s0 += PD) {
#pragma unroll
      for (int j = 0; j < PD; ++j) {
        const int s = s0 + j;
        if (s < T) {
          const int4 a = a_pre[j]; const int d = d_pre[j];
          if (s + PD < T) { a_pre[j] = __ldg(act + (int64_t)(s + PD) * n + env); d_pre[j] = __ldg(dem + (int64_t)(s + PD + 1) * n + env); }
          acc += (float)(a.x + a.y + a.z + a.w + d);
          for (int i = 0; i < delay; ++i) acc = acc * 1.0001f + 0.5f;
          float4 v = make_float4(acc, acc, acc, acc);
          float4* base = reinterpret_cast<float4*>(out + ((int64_t)s * n + warp_env0) * 28);
#pragma unroll
          for (int k = 0; k < 7; ++k) base[k * 32 + lane] = v;
        }
      }
    }
  }
}

__global__ void __launch_bounds__(64) k_slab(float* __restrict__ out, int64_t n, int T, int variant, int delay,
                                              int skew_ns) {
  const int64_t env = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (env >= n) return;
  // phase stagger: blocks that share an SM (b, b+148, b+296, ...) get different start delays so that their
  // store bursts do not coincide
  if (skew_ns > 0) {
    const unsigned phase = blockIdx.x % 7u;
    const long long t0 = clock64();
    const long long wait = (long long)phase * skew_ns * 2;       // ~2 cycles per ns at 1.97 GHz
    while (clock64() - t0 < wait) {}
  }
  const int lane = threadIdx.x & 31;
  const int64_t warp_env0 = env - lane;
  float acc = (float)env;
  for (int s = 0; s < T; ++s) {
    // some dependent arithmetic per period, like the recurrence (delay = number of dependent FMAs)
    for (int i = 0; i < delay; ++i) acc = acc * 1.0001f + 0.5f;
    float4 v = make_float4(acc, acc, acc, acc);
    float4* base;
    if (variant == 1) base = reinterpret_cast<float4*>(out + ((warp_env0 / 32) * T + s) * (32 * 28));
    else base = reinterpret_cast<float4*>(out + ((int64_t)s * n + warp_env0) * 28);
    // warp writes its 3,584 contiguous bytes as 7 fully coalesced 512-byte stores
    if (variant != 2) {
#pragma unroll
      for (int j = 0; j < 7; ++j) base[j * 32 + lane] = v;
    }
  }
  if (variant == 2) out[env] = acc;            // compute only: keep the chain alive
}

extern "C" int probe_write_in(float* out, const void* act, const void* dem, int64_t n, int T, int variant, int delay,
                              void* stream) {
  k_slab_in<<<(unsigned)((n + 63) / 64), 64, 0, (cudaStream_t)stream>>>(out, (const int4*)act, (const int*)dem, n, T, variant, delay);
  return (int)cudaGetLastError();
}

extern "C" int probe_write(float* out, int64_t n, int T, int variant, int delay, int skew_ns, void* stream) {
  k_slab<<<(unsigned)((n + 63) / 64), 64, 0, (cudaStream_t)stream>>>(out, n, T, variant, delay, skew_ns);
  return (int)cudaGetLastError();
}
