// random_fill.cu — device-side generation of demand / initial-state tables for THROUGHPUT runs
// (episode_source="device" of the VecEnv).  Philox-4x32-10 counter-based generator: element i of a
// fill with (seed, offset) is a pure function of (seed, offset + i/4), so tables are reproducible and
// independent of the launch geometry.  These streams have the reference's distributions
// (utils/demands.py:44,47; network_components.py:143,157,297) but are NOT numpy's MT19937 streams:
// parity runs use host-generated tables (utils/demands.py seeded_episode_inputs) instead.
#include "../../include/supplynet.h"
#include "supplynet_error.h"
#include "philox.cuh"

#include <cuda_runtime.h>
#include <stdint.h>

namespace snr {


template <typename T>
__global__ void __launch_bounds__(256)
k_fill_uniform(T* __restrict__ dst, int64_t count, int32_t lo, uint32_t range, uint64_t seed, uint64_t offset,
               const uint64_t* __restrict__ counter) {
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;        // quad index
  const int64_t i = q * 4;
  if (i >= count) return;
  const uint64_t c = offset + (uint64_t)q + (counter != nullptr ? *counter : 0ull);   // device-side stream position
  const uint4 r = philox4x32_10(make_uint4((uint32_t)c, (uint32_t)(c >> 32), 0u, 0u),
                                make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  const uint32_t u[4] = {r.x, r.y, r.z, r.w};
  T v[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) v[j] = (T)(lo + (int32_t)(((uint64_t)u[j] * range) >> 32));   // multiply-shift: bias < range / 2^32
  if (i + 4 <= count && sizeof(T) == 4) {
    *reinterpret_cast<uint4*>(dst + i) = *reinterpret_cast<const uint4*>(v);
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (i + j < count) dst[i + j] = v[j];
  }
}

// round(max(mean + sd * z, 0)) with z ~ N(0,1) by Box-Muller (utils/demands.py:47)
template <typename T>
__global__ void __launch_bounds__(256)
k_fill_normal_demand(T* __restrict__ dst, int64_t count, float mean, float sd, uint64_t seed, uint64_t offset,
                     const uint64_t* __restrict__ counter) {
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t i = q * 4;
  if (i >= count) return;
  const uint64_t c = offset + (uint64_t)q + (counter != nullptr ? *counter : 0ull);
  const uint4 r = philox4x32_10(make_uint4((uint32_t)c, (uint32_t)(c >> 32), 1u, 0u),
                                make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  const float k = 2.3283064365386963e-10f;                   // 2^-32
  const float u0 = ((float)r.x + 0.5f) * k, u1 = (float)r.y * k, u2 = ((float)r.z + 0.5f) * k, u3 = (float)r.w * k;
  const float r0 = sqrtf(-2.0f * logf(u0)), r1 = sqrtf(-2.0f * logf(u2));
  float s0, c0, s1, c1;
  sincospif(2.0f * u1, &s0, &c0);
  sincospif(2.0f * u3, &s1, &c1);
  const float z[4] = {r0 * c0, r0 * s0, r1 * c1, r1 * s1};
#pragma unroll
  for (int j = 0; j < 4; ++j)
    if (i + j < count) dst[i + j] = (T)rintf(fmaxf(mean + sd * z[j], 0.0f));
}

__global__ void k_counter_add(uint64_t* counter, uint64_t delta) { *counter += delta; }

// Row-major per-env table [n][width] (what the reference's reset produces per env: init values, demand stream) ->
// structure-of-arrays rows [rows >= width][ld] of the quantity type, converting on the way: 32 x 32 tiles through
// shared memory, both sides coalesced.  Columns n..ld-1 (row padding) and rows width..rows-1 are zero-filled.
template <typename S, typename T>
__global__ void __launch_bounds__(256)
k_pack_rows(const S* __restrict__ src, int64_t n, int width, int64_t src_stride, T* __restrict__ dst, int rows, int64_t ld,
            int64_t dst_row_stride) {
  __shared__ T tile[32][33];
  const int64_t env0 = (int64_t)blockIdx.x * 32;
  const int w0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;           // 32 x 8
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t e = env0 + ty + 8 * i;
    const int w = w0 + tx;
    tile[ty + 8 * i][tx] = (e < n && w < width) ? (T)src[e * src_stride + w] : T(0);
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int w = w0 + ty + 8 * i;
    const int64_t e = env0 + tx;
    if (w < rows && e < ld) dst[(int64_t)w * dst_row_stride + e] = tile[tx][ty + 8 * i];
  }
}

}  // namespace snr

extern "C" {

int32_t sn_fill_uniform_at(int32_t dtype, void* dst, int64_t count, int32_t low, int32_t high_exclusive, uint64_t seed,
                           const uint64_t* counter, uint64_t offset, void* stream) {
  using sn_detail::fail;
  if (!dst || count < 1 || high_exclusive <= low) return fail(SN_ERR_INVALID, "sn_fill_uniform: dst NULL, count < 1 or empty range");
  if ((reinterpret_cast<uintptr_t>(dst) & 15u) != 0) return fail(SN_ERR_INVALID, "sn_fill_uniform: dst must be 16-byte aligned");
  const uint32_t range = (uint32_t)(high_exclusive - low);
  const unsigned grid = (unsigned)(((count + 3) / 4 + 255) / 256);
  cudaStream_t st = (cudaStream_t)stream;
  switch (dtype) {
    case SN_I32: snr::k_fill_uniform<int32_t><<<grid, 256, 0, st>>>((int32_t*)dst, count, low, range, seed, offset, counter); break;
    case SN_F32: snr::k_fill_uniform<float><<<grid, 256, 0, st>>>((float*)dst, count, low, range, seed, offset, counter); break;
    case SN_F64: snr::k_fill_uniform<double><<<grid, 256, 0, st>>>((double*)dst, count, low, range, seed, offset, counter); break;
    default: return sn_detail::fail(SN_ERR_INVALID, "sn_fill: unknown dtype %d", dtype);
  }
  return cudaGetLastError() == cudaSuccess ? SN_OK : sn_detail::fail(SN_ERR_CUDA, "sn_fill: kernel launch failed");
}

int32_t sn_fill_uniform(int32_t dtype, void* dst, int64_t count, int32_t low, int32_t high_exclusive, uint64_t seed,
                        uint64_t offset, void* stream) {
  return sn_fill_uniform_at(dtype, dst, count, low, high_exclusive, seed, nullptr, offset, stream);
}

int32_t sn_fill_normal_demand_at(int32_t dtype, void* dst, int64_t count, double mean, double sd, uint64_t seed,
                                 const uint64_t* counter, uint64_t offset, void* stream) {
  using sn_detail::fail;
  if (!dst || count < 1 || sd < 0) return fail(SN_ERR_INVALID, "sn_fill_normal_demand: dst NULL, count < 1 or negative sd");
  const unsigned grid = (unsigned)(((count + 3) / 4 + 255) / 256);
  cudaStream_t st = (cudaStream_t)stream;
  switch (dtype) {
    case SN_I32: snr::k_fill_normal_demand<int32_t><<<grid, 256, 0, st>>>((int32_t*)dst, count, (float)mean, (float)sd, seed, offset, counter); break;
    case SN_F32: snr::k_fill_normal_demand<float><<<grid, 256, 0, st>>>((float*)dst, count, (float)mean, (float)sd, seed, offset, counter); break;
    case SN_F64: snr::k_fill_normal_demand<double><<<grid, 256, 0, st>>>((double*)dst, count, (float)mean, (float)sd, seed, offset, counter); break;
    default: return sn_detail::fail(SN_ERR_INVALID, "sn_fill: unknown dtype %d", dtype);
  }
  return cudaGetLastError() == cudaSuccess ? SN_OK : sn_detail::fail(SN_ERR_CUDA, "sn_fill: kernel launch failed");
}

int32_t sn_fill_normal_demand(int32_t dtype, void* dst, int64_t count, double mean, double sd, uint64_t seed,
                              uint64_t offset, void* stream) {
  return sn_fill_normal_demand_at(dtype, dst, count, mean, sd, seed, nullptr, offset, stream);
}

int32_t sn_counter_add(uint64_t* counter, uint64_t delta, void* stream) {
  if (!counter) return sn_detail::fail(SN_ERR_INVALID, "sn_counter_add: counter is NULL");
  snr::k_counter_add<<<1, 1, 0, (cudaStream_t)stream>>>(counter, delta);
  return cudaGetLastError() == cudaSuccess ? SN_OK : sn_detail::fail(SN_ERR_CUDA, "sn_counter_add: kernel launch failed");
}

int32_t sn_pack_rows(int32_t src_dtype, const void* src, int64_t n_envs, int32_t width, int64_t src_stride, int32_t dtype,
                     void* dst, int32_t rows, int64_t ld, int64_t dst_row_stride, void* stream) {
  using sn_detail::fail;
  if (!src || !dst) return fail(SN_ERR_INVALID, "sn_pack_rows: NULL argument");
  if (n_envs < 1 || width < 1 || rows < width || ld < n_envs || src_stride < width || dst_row_stride < ld)
    return fail(SN_ERR_INVALID, "sn_pack_rows: n_envs=%lld width=%d rows=%d ld=%lld", (long long)n_envs, width, rows, (long long)ld);
  const dim3 grid((unsigned)((ld + 31) / 32), (unsigned)((rows + 31) / 32));
  cudaStream_t st = (cudaStream_t)stream;
#define SN_PACK(S, T) snr::k_pack_rows<S, T><<<grid, 256, 0, st>>>((const S*)src, n_envs, width, src_stride, (T*)dst, rows, ld, dst_row_stride)
#define SN_PACK_TO(S)                                                             \
  switch (dtype) {                                                               \
    case SN_I32: SN_PACK(S, int32_t); break;                                     \
    case SN_F32: SN_PACK(S, float); break;                                       \
    case SN_F64: SN_PACK(S, double); break;                                      \
    default: return fail(SN_ERR_INVALID, "sn_pack_rows: unknown dtype %d", dtype); \
  }
  switch (src_dtype) {
    case SN_SRC_I32: SN_PACK_TO(int32_t); break;
    case SN_SRC_I64: SN_PACK_TO(int64_t); break;
    case SN_SRC_F32: SN_PACK_TO(float); break;
    case SN_SRC_F64: SN_PACK_TO(double); break;
    default: return fail(SN_ERR_INVALID, "sn_pack_rows: unknown source dtype %d", src_dtype);
  }
#undef SN_PACK_TO
#undef SN_PACK
  return cudaGetLastError() == cudaSuccess ? SN_OK : fail(SN_ERR_CUDA, "sn_pack_rows: kernel launch failed");
}

}  // extern "C"
