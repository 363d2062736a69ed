// supplynet_kernels.cu — sm_100a kernels and the C ABI of include/supplynet.h.
//
// Kernels (one thread = one env, structure-of-arrays state so every row access of a warp is
// one fully used 128-byte line):
//   k_reset    init rows -> state rows (+ observation)
//   k_step     one period for all envs: 40 state rows in, 40 out, actions/demand in, reward out
//   k_observe  state -> observation
//   k_rollout  n_steps periods with the env held in registers; actions streamed or base-stock
//              policy evaluated in-kernel; optional trajectory out
//   k_heuristic  BaseStockPolicy on a batch of observations
// Every env kernel is written against an `Ops` trait: StdOps = the beer game's lead times (tuned,
// supplynet_core.cuh), GenOps = run-time per-arc lead times and shorter chains (supplynet_generic.cuh),
// TreeOps = distribution trees of up to four nodes (supplynet_tree.cuh).  Trees of five to eight nodes have
// their own kernels with one lane per node, five to eight lanes per env (k_net_*, supplynet_net.cuh).  The device-side VecNormalize
// kernels live in vecnorm_kernels.cu.
// Row-major float observations ([n_envs][obs_dim], what a VecEnv returns) are staged per warp
// in shared memory and written with one TMA bulk store (cp.async.bulk.global.shared::cta) so
// the 112-byte-per-thread pattern reaches L2 as full lines.
#include "../../include/supplynet.h"
#include "supplynet_error.h"
#include "supplynet_core.cuh"
#include "supplynet_tree.cuh"
#include "supplynet_net.cuh"

#include <cuda.h>              // CUtensorMap (types only: the encoder is fetched through cudaGetDriverEntryPoint)

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <type_traits>

namespace sn {

#ifndef SN_BLOCK
#define SN_BLOCK 128
#endif
#ifndef SN_STEP_MIN_BLOCKS
#define SN_STEP_MIN_BLOCKS 1
#endif
constexpr int kBlock = SN_BLOCK;
constexpr int kWarpsPerBlock = kBlock / 32;

// ------------------------------------------------------------------------------------------
// small PTX helpers
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void bulk_store_s2g(void* gdst, const void* ssrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst),
               "r"((uint32_t)__cvta_generic_to_shared(ssrc)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }


// ------------------------------------------------------------------------------------------
// Observation tile writer.  Each warp owns 32 consecutive envs = one contiguous span of
// 32*od floats in the row-major output.  Lanes deposit their `od` floats in shared memory
// (stride od words: conflict-free for od = 7, 21 and for od = 28 written as float4), then lane 0
// issues one bulk store.  Partial warps (tail of the batch) and unaligned outputs use plain stores.
constexpr int kObsMax = 7 * kF;       // floats of the widest observation row (28; 56 in eight-node builds)
struct ObsStager {
  float* smem;     // this warp's staging area: [2][32*kObsMax] floats
  int lane;
  __device__ __forceinline__ ObsStager(float* block_smem) {
    lane = threadIdx.x & 31;
    smem = block_smem + (threadIdx.x >> 5) * (2 * 32 * kObsMax);
  }
  // Staging row of this lane in buffer `buf`.  `buf` must alternate 0/1 between consecutive stores of the same
  // warp (at most one earlier bulk store, the one that used the other buffer, may still be reading shared memory);
  // a caller that cannot guarantee the alternation drains first.
  __device__ __forceinline__ float* begin(int od, int buf) {
    if (lane == 0) bulk_wait_read<1>();
    __syncwarp();
    return smem + buf * (32 * kObsMax) + lane * od;
  }
  __device__ __forceinline__ void commit(float* gdst_warp, int od, int buf) {
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
      bulk_store_s2g(gdst_warp, smem + buf * (32 * kObsMax), (uint32_t)(32 * od * sizeof(float)));
      bulk_commit();
    }
  }
  __device__ __forceinline__ void drain() {
    if (lane == 0) bulk_wait_read<0>();
    __syncwarp();
  }
};

// Writes this env's observation row: through the warp's TMA staging buffer when the warp is full
// and the output aligned, with plain stores otherwise.
template <bool IDENT, typename T, typename E>
__device__ __forceinline__ void write_obs(ObsStager& st, const E& e, const DevSpec<T>& sp, bool terminal,
                                          float* __restrict__ gobs, int64_t env, int64_t n, int od, bool valid,
                                          int buf, bool use_tma) {
  const int64_t env0 = env - st.lane;
  if (use_tma && env0 + 32 <= n) {
    float* row = st.begin(od, buf);
    emit_obs<IDENT>(e, sp, terminal, row);
    st.commit(gobs + env0 * od, od, buf);
  } else if (valid) {
    emit_obs<IDENT>(e, sp, terminal, gobs + env * od);
  }
}

// actions are row-major [n][na]; a[k] receives the action of NODE k (co-players: untouched)
template <bool IDENT, typename T>
__device__ __forceinline__ void load_actions(const T* __restrict__ act, int64_t env, const DevSpec<T>& sp, T (&a)[kF]) {
  if (IDENT) {
    if (sizeof(T) == 4) {
      const int4 v = __ldg(reinterpret_cast<const int4*>(act + env * 4));
      a[0] = *reinterpret_cast<const T*>(&v.x);
      a[1] = *reinterpret_cast<const T*>(&v.y);
      a[2] = *reinterpret_cast<const T*>(&v.z);
      a[3] = *reinterpret_cast<const T*>(&v.w);
    } else {
      const int4 lo = __ldg(reinterpret_cast<const int4*>(act + env * 4));
      const int4 hi = __ldg(reinterpret_cast<const int4*>(act + env * 4) + 1);
      const T* pl = reinterpret_cast<const T*>(&lo);
      const T* ph = reinterpret_cast<const T*>(&hi);
      a[0] = pl[0]; a[1] = pl[1]; a[2] = ph[0]; a[3] = ph[1];
    }
  } else {
    const T* p = act + env * sp.n_agents;
#pragma unroll
    for (int k = 0; k < kF; ++k)
      if (sp.slot[k] >= 0) a[k] = __ldg(p + sp.slot[k]);
  }
}

template <bool IDENT, typename T>
__device__ __forceinline__ void store_actions_f32(float* __restrict__ dst, int64_t env, const DevSpec<T>& sp,
                                                  const T (&a)[kF]) {
  if (IDENT) {
    *reinterpret_cast<float4*>(dst + env * 4) = make_float4((float)a[0], (float)a[1], (float)a[2], (float)a[3]);
  } else {
    float* p = dst + env * sp.n_agents;
#pragma unroll
    for (int k = 0; k < kF; ++k)
      if (sp.slot[k] >= 0) p[sp.slot[k]] = (float)a[k];
  }
}

// the same from a shared-memory tile of action rows ([tile][n_agents], row `row`)
template <bool IDENT, typename T>
__device__ __forceinline__ void load_actions_smem(const T* s_act, int row, const DevSpec<T>& sp, T (&a)[kF]) {
  if (IDENT) {
    if (sizeof(T) == 4) {
      const int4 v = *reinterpret_cast<const int4*>(s_act + row * 4);
      a[0] = *reinterpret_cast<const T*>(&v.x);
      a[1] = *reinterpret_cast<const T*>(&v.y);
      a[2] = *reinterpret_cast<const T*>(&v.z);
      a[3] = *reinterpret_cast<const T*>(&v.w);
    } else {
#pragma unroll
      for (int k = 0; k < kF; ++k) a[k] = s_act[row * 4 + k];
    }
  } else {
    const T* p = s_act + row * sp.n_agents;
#pragma unroll
    for (int k = 0; k < kF; ++k)
      if (sp.slot[k] >= 0) a[k] = p[sp.slot[k]];
  }
}

// block-level reduction of per-thread totals into stats[] (one atomic per block per word)
template <int NV, int WARPS>
__device__ __forceinline__ void block_reduce_add(double (&v)[NV], double* __restrict__ stats, double* red /*[WARPS][NV]*/) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    double x = v[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if (lane == 0) red[warp * NV + i] = x;
  }
  __syncthreads();
  if (threadIdx.x < NV) {
    double x = 0.0;
#pragma unroll
    for (int w = 0; w < WARPS; ++w) x += red[w * NV + threadIdx.x];
    atomicAdd(&stats[threadIdx.x], x);
  }
}

// Per-topology operations the kernels are written against: StdOps = the beer game's lead times
// (tuned, 40 state words), GenOps = run-time per-arc lead times (supplynet_generic.cuh, 57 words).
template <typename T>
struct StdOps {
  static constexpr bool kWide = false;
  static constexpr int kWords = kStateWords;       // rows of the state buffer
  static constexpr int kMaxSrc = 1;                // demand rows per period
  using E = Env<T>;
  using D = T;                                     // demand of one period: one number per env
  static __device__ __forceinline__ int64_t dstride(const DevSpec<T>&, int64_t ld) { return ld; }
  static __device__ __forceinline__ int n_src(const DevSpec<T>&) { return 1; }
  static __device__ __forceinline__ D load_demand_smem(const T* row, int, int col, const DevSpec<T>&) { return row[col]; }
  static __device__ __forceinline__ D load_demand(const T* __restrict__ row, int64_t, int64_t env, const DevSpec<T>&) {
    return __ldg(row + env);
  }
  static __device__ __forceinline__ void load(const DevSpec<T>&, const T* __restrict__ state, int64_t ld, int64_t env, E& e) {
    T w[kStateWords];
#pragma unroll
    for (int i = 0; i < kStateWords; ++i) w[i] = state[i * ld + env];
    env_from_words(e, w);
  }
  static __device__ __forceinline__ void store(T* __restrict__ state, int64_t ld, int64_t env, const E& e) {
    T w[kStateWords];
    env_to_words(e, w);
#pragma unroll
    for (int i = 0; i < kStateWords; ++i) state[i * ld + env] = w[i];
  }
  static __device__ __forceinline__ void reset(E& e, const DevSpec<T>&, const T* __restrict__ init, int64_t ld,
                                               int64_t env, T d0) {
    T in[kInitWords];
#pragma unroll
    for (int w = 0; w < kInitWords; ++w) in[w] = __ldg(init + w * ld + env);
    env_reset(e, in, d0);
  }
  template <bool MULTI>
  static __device__ __forceinline__ void step(E& e, const DevSpec<T>& sp, int item, const T (&q)[kF], T dn, bool terminal,
                                              StepOut& so) {
    env_step<MULTI>(e, sp, item, q, dn, terminal, so);
  }
};

template <typename T>
struct GenOps {
  static constexpr bool kWide = false;            // measured: 0.209 ms with 128 registers + 14 spill accesses, 0.245 ms wide
  static constexpr int kWords = kGStateWords;
  static constexpr int kMaxSrc = 1;
  using E = GEnv<T>;
  using D = T;
  static __device__ __forceinline__ int64_t dstride(const DevSpec<T>&, int64_t ld) { return ld; }
  static __device__ __forceinline__ int n_src(const DevSpec<T>&) { return 1; }
  static __device__ __forceinline__ D load_demand_smem(const T* row, int, int col, const DevSpec<T>&) { return row[col]; }
  static __device__ __forceinline__ D load_demand(const T* __restrict__ row, int64_t, int64_t env, const DevSpec<T>&) {
    return __ldg(row + env);
  }
  static __device__ __forceinline__ void load(const DevSpec<T>&, const T* __restrict__ state, int64_t ld, int64_t env, E& e) {
    T w[kGStateWords];
#pragma unroll
    for (int i = 0; i < kGStateWords; ++i) w[i] = state[i * ld + env];
    genv_from_words(e, w);
  }
  static __device__ __forceinline__ void store(T* __restrict__ state, int64_t ld, int64_t env, const E& e) {
    T w[kGStateWords];
    genv_to_words(e, w);
#pragma unroll
    for (int i = 0; i < kGStateWords; ++i) state[i * ld + env] = w[i];
  }
  static __device__ __forceinline__ void reset(E& e, const DevSpec<T>& sp, const T* __restrict__ init, int64_t ld,
                                               int64_t env, T d0) {
    T in[kGInitWords];
#pragma unroll
    for (int w = 0; w < kGInitWords; ++w) in[w] = __ldg(init + w * ld + env);
    genv_reset(e, sp, in, d0);
  }
  template <bool MULTI>
  static __device__ __forceinline__ void step(E& e, const DevSpec<T>& sp, int item, const T (&q)[kF], T dn, bool terminal,
                                              StepOut& so) {
    genv_step(e, sp, MULTI ? item : 0, q, dn, terminal, so);
  }
};

// distribution trees (supplynet_tree.cuh, 64 words): one demand row per source and period
template <typename T>
struct TreeOps {
#if defined(SN_STATIC_TOPO) && !defined(SN_ST_WIDE)
  static constexpr bool kWide = false;            // specialised builds fit the 128 registers of the single-wave geometry
#else
  static constexpr bool kWide = true;             // k_rollout_wide, 144 registers: 0.315 ms vs 0.355 ms with 128 + ~100 spill accesses
#endif
  static constexpr int kWords = kTStateWords;
  static constexpr int kMaxSrc = kF;               // one demand row per source
  using E = TEnv<T>;
  using D = TreeDem<T>;
  static __device__ __forceinline__ int64_t dstride(const DevSpec<T>& sp, int64_t ld) { return ld * sp.n_src; }
  static __device__ __forceinline__ int n_src(const DevSpec<T>& sp) { return sp.n_src; }
  static __device__ __forceinline__ D load_demand_smem(const T* rows, int tile, int col, const DevSpec<T>& sp) {
    D d;
#pragma unroll
    for (int k = 0; k < kF; ++k) d.v[k] = sp.src[k] >= 0 ? rows[sp.src[k] * tile + col] : T(0);
    return d;
  }
  static __device__ __forceinline__ D load_demand(const T* __restrict__ row, int64_t ld, int64_t env, const DevSpec<T>& sp) {
    D d;
#pragma unroll
    for (int k = 0; k < kF; ++k) d.v[k] = sp.src[k] >= 0 ? __ldg(row + sp.src[k] * ld + env) : T(0);
    return d;
  }
  static __device__ __forceinline__ void load(const DevSpec<T>& sp, const T* __restrict__ state, int64_t ld, int64_t env, E& e) {
    T w[kTStateWords];
#pragma unroll
    for (int i = 0; i < kTStateWords; ++i) w[i] = state[i * ld + env];
    tenv_from_words(e, w);
    tenv_static_zeros(e);
  }
  static __device__ __forceinline__ void store(T* __restrict__ state, int64_t ld, int64_t env, const E& e) {
    T w[kTStateWords];
    tenv_to_words(e, w);
#pragma unroll
    for (int i = 0; i < kTStateWords; ++i) state[i * ld + env] = w[i];
  }
  static __device__ __forceinline__ void reset(E& e, const DevSpec<T>& sp, const T* __restrict__ init, int64_t ld,
                                               int64_t env, const D& d0) {
    T in[kGInitWords];
#pragma unroll
    for (int w = 0; w < kGInitWords; ++w) in[w] = __ldg(init + w * ld + env);
    tenv_reset(e, sp, in, d0);
    tenv_static_zeros(e);
  }
  template <bool MULTI>
  static __device__ __forceinline__ void step(E& e, const DevSpec<T>& sp, int, const T (&q)[kF], const D& dn, bool terminal,
                                              StepOut& so) {
    tenv_step(e, sp, q, dn, terminal, so);
  }
};

// ------------------------------------------------------------------------------------------
template <typename T, bool IDENT, typename Ops>
__global__ void __launch_bounds__(kBlock)
k_reset(const __grid_constant__ DevSpec<T> sp, int64_t n, int64_t ld, const T* __restrict__ init,
        const T* __restrict__ demand0, T* __restrict__ state, float* __restrict__ obs, int use_tma) {
  extern __shared__ __align__(16) float smem_f[];
  const int64_t env = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  const bool valid = env < n;
  typename Ops::E e;
  if (valid) {
    Ops::reset(e, sp, init, ld, env, Ops::load_demand(demand0, ld, env, sp));
    Ops::store(state, ld, env, e);
  }
  if (obs != nullptr) {
    ObsStager st(smem_f);
    write_obs<IDENT>(st, e, sp, false, obs, env, n, 7 * sp.n_obs, valid, 0, use_tma != 0);
    st.drain();
  }
}

template <typename T, bool IDENT, typename Ops>
__global__ void __launch_bounds__(kBlock)
k_observe(const __grid_constant__ DevSpec<T> sp, int64_t n, int64_t ld, const T* __restrict__ state,
          float* __restrict__ obs, int use_tma) {
  extern __shared__ __align__(16) float smem_f[];
  const int64_t env = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  const bool valid = env < n;
  typename Ops::E e;
  if (valid) Ops::load(sp, state, ld, env, e);
  ObsStager st(smem_f);
  write_obs<IDENT>(st, e, sp, false, obs, env, n, 7 * sp.n_obs, valid, 0, use_tma != 0);
  st.drain();
}

struct StepPtrs {
  float* obs;
  float* reward;
  double* reward64;
  double* cost;
  double* ep_return;
  double* stats;
  float* env_reward;        // multi-item batches: reward of env = sum over its n_items adjacent chains, [n/n_items]
  int32_t* ctl;             // device control words (SN_CTL_*) or NULL: period read from / advanced in device memory
  // fused VecNormalize statistics (tiled kernel only; all NULL = off)
  double* vn_scratch;       // [2*od + 2], zero between launches
  double* vn_obs_rms;       // mean[od], var[od], count  (updated in place by the last block)
  double* vn_ret_rms;       // mean, var, count
  double* vn_returns;       // [n] discounted returns (NULL = observations only)
  double vn_gamma;
  // one-kernel mode (tiled kernel, at most one tile per SM): normalised outputs written by the same launch
  float* vn_nobs;           // [n][od]
  float* vn_nrew;           // [n] (NULL = rewards are not normalised)
  float vn_clip_obs, vn_clip_rew;
  double vn_eps;
};

// Period of this launch: the host's argument, or the device counter ctl[SN_CTL_PERIOD] (one captured CUDA graph
// then serves every period of an episode).  The counter is advanced by the launch itself, by whichever block
// finishes last (ticket in ctl[SN_CTL_TICKET]); every block reads it before it takes its ticket.
__device__ __forceinline__ int launch_period(const int32_t* ctl, int period_arg) {
  return ctl != nullptr ? *reinterpret_cast<const volatile int32_t*>(ctl + SN_CTL_PERIOD) : period_arg;
}
// true for exactly one thread of the grid: thread 0 of the block that finishes last
__device__ __forceinline__ bool last_block_ticket(int32_t* ctl) {
  __threadfence();
  return atomicAdd(ctl + SN_CTL_TICKET, 1) == (int)gridDim.x - 1;
}

// Sum of the step rewards of the n_items adjacent chains of one env (item-minor batches, n_items 2 or 4: an
// env never straddles a warp), in item order, in double.  Every lane of the warp must call; the lane of item 0
// gets the total.
__device__ __forceinline__ double env_reward_sum(double r, int n_items) {
  const int lane = threadIdx.x & 31;
  if (n_items == 2) {
    const double o = __shfl_xor_sync(0xffffffffu, r, 1);
    return (lane & 1) ? o + r : r + o;
  }
  const int base = lane & ~3;
  const double r0 = __shfl_sync(0xffffffffu, r, base), r1 = __shfl_sync(0xffffffffu, r, base + 1);
  const double r2 = __shfl_sync(0xffffffffu, r, base + 2), r3 = __shfl_sync(0xffffffffu, r, base + 3);
  return ((r0 + r1) + r2) + r3;
}

template <typename T, bool IDENT, typename Ops>
__global__ void __launch_bounds__(kBlock, SN_STEP_MIN_BLOCKS)
k_step(const __grid_constant__ DevSpec<T> sp, int64_t n, int64_t ld, int period_arg, T* __restrict__ state,
       const T* __restrict__ actions, const T* __restrict__ demand, const StepPtrs out, int use_tma) {
  extern __shared__ __align__(16) float smem_f[];
  __shared__ double red[kWarpsPerBlock * 14];
  const int period = launch_period(out.ctl, period_arg);
  if (period >= sp.t_max) {                    // only reachable through the device counter (the host checks its own)
    if (out.ctl != nullptr && blockIdx.x == 0 && threadIdx.x == 0) atomicOr(out.ctl + SN_CTL_FLAGS, SN_FLAG_STEP_AFTER_TERMINAL);
    return;
  }
  const bool terminal = period + 1 >= sp.t_max;
  // with a device counter `demand` is the whole table and the row of period+1 is picked here
  const T* __restrict__ demand_next = out.ctl != nullptr ? demand + (int64_t)(period + 1) * Ops::dstride(sp, ld) : demand;
  const int64_t env = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  const bool valid = env < n;
  typename Ops::E e;
  StepOut so;
  so.reward = 0.0;
  bool over = false;
#pragma unroll
  for (int k = 0; k < kF; ++k) so.hold[k] = so.back[k] = 0.0;
  if (valid) {
    Ops::load(sp, state, ld, env, e);
    T a[kF] = {T(0), T(0), T(0), T(0)};
    load_actions<IDENT>(actions, env, sp, a);
    const typename Ops::D dn = Ops::load_demand(demand_next, ld, env, sp);
    T q[kF];
    env_orders<IDENT>(e, sp, a, q);
    const uint32_t om = qty_mask_orders(a, q);
    if (sp.n_items > 1) Ops::template step<true>(e, sp, chain_item(sp, env), q, dn, terminal, so);
    else Ops::template step<false>(e, sp, 0, q, dn, terminal, so);
    over = qty_mask_over(om | qty_mask_state<T>(e));
    Ops::store(state, ld, env, e);
    if (out.reward) out.reward[env] = (float)so.reward;
    if (out.reward64) out.reward64[env] = so.reward;
    if (out.cost) {
#pragma unroll
      for (int k = 0; k < kF; ++k) {
        out.cost[k * ld + env] = so.hold[k];
        out.cost[(kF + k) * ld + env] = so.back[k];
      }
    }
    if (out.ep_return) out.ep_return[env] += so.reward;
  }
  if (out.env_reward != nullptr) {             // uniform branch: every lane takes part in the shuffles
    const double tot = env_reward_sum(so.reward, sp.n_items);
    if (valid && (env % sp.n_items) == 0) out.env_reward[env / sp.n_items] = (float)tot;
  }
  if (out.obs != nullptr) {
    ObsStager st(smem_f);
    write_obs<IDENT>(st, e, sp, terminal, out.obs, env, n, 7 * sp.n_obs, valid, 0, use_tma != 0);
    st.drain();
  }
  if (over && out.ctl != nullptr) atomicOr(out.ctl + SN_CTL_FLAGS, SN_FLAG_QTY_RANGE);
  if (out.stats != nullptr) {
    double v[14];
    v[0] = valid ? 1.0 : 0.0;
    v[1] = so.reward;
#pragma unroll
    for (int k = 0; k < 4; ++k) {          // per-facility sums: four-node builds only (like the lane-per-node kernels)
      v[2 + k] = kF == 4 ? so.hold[k] : 0.0;
      v[6 + k] = kF == 4 ? so.back[k] : 0.0;
    }
    // episode moments (Monitor's episode statistics) when this step ends the episode
    const bool ep = valid && terminal && out.ep_return != nullptr;
    const double ret = ep ? out.ep_return[env] : 0.0;
    v[10] = ep ? 1.0 : 0.0;
    v[11] = ret;
    v[12] = ret * ret;
    v[13] = over ? 1.0 : 0.0;
    block_reduce_add<14, kWarpsPerBlock>(v, out.stats, red);
  }
  if (out.ctl != nullptr) {
    __syncthreads();
    if (threadIdx.x == 0 && last_block_ticket(out.ctl)) {
      out.ctl[SN_CTL_TICKET] = 0;
      out.ctl[SN_CTL_PERIOD] = period + 1;
    }
  }
}

#include "supplynet_step_tiled.cuh"

struct RolloutPtrs {
  float* traj_obs;
  float* traj_reward;
  float* traj_actions;
  float* obs;
  double* ep_return;
  double* stats;
};

#ifndef SN_ROLLOUT_BLOCK
#define SN_ROLLOUT_BLOCK 64             // small blocks: 65,536 envs spread evenly over 148 SMs
#endif
#ifndef SN_ROLLOUT_UNROLL
#define SN_ROLLOUT_UNROLL 1             // A/B on B200: 1..4 within noise for the trajectory rollout, 1 best for the policy rollout
#endif
#define SN_STR2(x) #x
#define SN_STR(x) SN_STR2(x)
#define SN_PRAGMA_UNROLL(n) _Pragma(SN_STR(unroll n))
constexpr int kRolloutBlock = SN_ROLLOUT_BLOCK;

// body of the fused rollout kernels below
template <typename T, int POLICY, bool IDENT, bool MULTI, typename Ops>
__device__ __forceinline__ void
rollout_body(const DevSpec<T>& sp, const DevPolicy<T>& pol, int64_t n, int64_t ld,
             int period0, int n_steps, T* __restrict__ state, const T* __restrict__ demand,
             const T* __restrict__ actions, const RolloutPtrs& out, int use_tma, const T* __restrict__ init,
             float* __restrict__ obs0, float* smem_f, double* red) {
  const int64_t env = (int64_t)blockIdx.x * kRolloutBlock + threadIdx.x;
  const bool valid = env < n;
  const int64_t envc = valid ? env : (n - 1);       // tail lanes shadow the last env (loads only)
  const int od = 7 * sp.n_obs;
  const int64_t act_stride = n * sp.n_agents;
  ObsStager st(smem_f);

  typename Ops::E e;
  if (init != nullptr) {
    // episode mode (sn_episode): reset in-kernel, the state never has to exist in HBM
    Ops::reset(e, sp, init, ld, envc, Ops::load_demand(demand, ld, envc, sp));
    if (obs0 != nullptr) write_obs<IDENT>(st, e, sp, false, obs0, env, n, od, valid, 1, use_tma != 0);
  } else {
    Ops::load(sp, state, ld, envc, e);
  }
  double acc[10];
#pragma unroll
  for (int i = 0; i < 10; ++i) acc[i] = 0.0;
  bool qover = false;                               // range guard of the integer path (supplynet_core.cuh); a predicate, not a register
  const int item = MULTI ? chain_item(sp, envc) : 0;

  // software pipeline: inputs of step s+1 are requested before step s is computed
  T a_pre[kF] = {T(0), T(0), T(0), T(0)};
  if (POLICY == SN_POLICY_ACTIONS) load_actions<IDENT>(actions, envc, sp, a_pre);
  const int64_t dstride = Ops::dstride(sp, ld);
  typename Ops::D d_pre = Ops::load_demand(demand + (int64_t)(period0 + 1) * dstride, ld, envc, sp);

SN_PRAGMA_UNROLL(SN_ROLLOUT_UNROLL)
  for (int s = 0; s < n_steps; ++s) {
    const int t = period0 + s;
    const bool terminal = (t + 1 >= sp.t_max);
    T a[kF];
#pragma unroll
    for (int i = 0; i < kF; ++i) a[i] = a_pre[i];
    const typename Ops::D dn = d_pre;
    if (s + 1 < n_steps) {
      if (POLICY == SN_POLICY_ACTIONS) load_actions<IDENT>(actions + (int64_t)(s + 1) * act_stride, envc, sp, a_pre);
      d_pre = Ops::load_demand(demand + (int64_t)(t + 2) * dstride, ld, envc, sp);
    }
    if (POLICY == SN_POLICY_BASE_STOCK) {
      // BaseStockPolicy.get_order_quantity on this env's own observation (utils/heuristics.py:43-57)
#pragma unroll
      for (int k = 0; k < kF; ++k)
        if (IDENT || sp.slot[k] >= 0)
          a[k] = base_stock(pol.target[k], e.inv[k], e.U[k], e.unfilled(k, sp), e.latest(k), pol.lb, pol.ub, pol.rule);
    }
    if (out.traj_actions != nullptr && valid) store_actions_f32<IDENT>(out.traj_actions + (int64_t)s * act_stride, env, sp, a);
    T q[kF];
    env_orders<IDENT>(e, sp, a, q);
    // (not in the tree rollouts: their register budget is spent, the guard's few registers cost them 45 % in
    // spills; sn_step guards every topology)
    if (POLICY == SN_POLICY_ACTIONS && !Ops::kWide) qover |= qty_mask_over(qty_mask_orders(a, q));
    StepOut so;
    Ops::template step<MULTI>(e, sp, item, q, dn, terminal, so);
    if (!Ops::kWide) qover |= qty_mask_over(qty_mask_state<T>(e));
    acc[0] += 1.0;
    acc[1] += so.reward;
    if (kF == 4) {
#pragma unroll
      for (int k = 0; k < 4; ++k) { acc[2 + k] += so.hold[k]; acc[6 + k] += so.back[k]; }
    }
    if (out.traj_reward != nullptr && valid) out.traj_reward[(int64_t)s * n + env] = (float)so.reward;
    if (out.traj_obs != nullptr)
      write_obs<IDENT>(st, e, sp, terminal, out.traj_obs + (int64_t)s * n * od, env, n, od, valid, s & 1, use_tma != 0);
  }

  const bool ended = (period0 + n_steps >= sp.t_max);
  if (out.obs != nullptr) {
    // without a trajectory nothing has alternated the staging buffers since the obs0 store (buffer 1): with an odd
    // n_steps this store would reuse that buffer while the bulk copy may still be reading it
    if (out.traj_obs == nullptr) st.drain();
    write_obs<IDENT>(st, e, sp, ended, out.obs, env, n, od, valid, n_steps & 1, use_tma != 0);
  }
  st.drain();
  double ep_total = acc[1];                 // return of the whole episode, if this launch can know it
  if (valid) {
    if (state != nullptr) Ops::store(state, ld, env, e);
    if (out.ep_return) {
      ep_total = (init != nullptr ? 0.0 : out.ep_return[env]) + acc[1];
      out.ep_return[env] = ep_total;
    }
  }
  if (out.stats != nullptr) {
    double v[14];
#pragma unroll
    for (int i = 0; i < 10; ++i) v[i] = valid ? acc[i] : 0.0;
    // episode moments when the episode ends in this launch and its full return is known here
    const bool ep = valid && ended && (init != nullptr || out.ep_return != nullptr || period0 == 0);
    v[10] = ep ? 1.0 : 0.0;
    v[11] = ep ? ep_total : 0.0;
    v[12] = ep ? ep_total * ep_total : 0.0;
    v[13] = (valid && qover) ? 1.0 : 0.0;                      // envs (not env-steps) that left the exact int32 range
    block_reduce_add<14, kRolloutBlock / 32>(v, out.stats, red);
  }
}

#define SN_ROLLOUT_PARAMS                                                                                          \
  const __grid_constant__ DevSpec<T> sp, const __grid_constant__ DevPolicy<T> pol, int64_t n, int64_t ld, int period0, \
      int n_steps, T *__restrict__ state, const T *__restrict__ demand, const T *__restrict__ actions,                 \
      const RolloutPtrs out, int use_tma, const T *__restrict__ init, float *__restrict__ obs0
#define SN_ROLLOUT_ARGS sp, pol, n, ld, period0, n_steps, state, demand, actions, out, use_tma, init, obs0

// tuned beer-game chain: 7 blocks of 64 threads per SM: 65,536 envs = 1,024 blocks fit the 148 SMs in ONE wave
#ifndef SN_ROLLOUT_MINB8
#define SN_ROLLOUT_MINB8 4              // eight-node specialised builds: up to 255 registers per thread
#endif
template <typename T, int POLICY, bool IDENT, bool MULTI, typename Ops>
__global__ void __launch_bounds__(kRolloutBlock, kF == 4 ? 448 / kRolloutBlock : SN_ROLLOUT_MINB8) k_rollout(SN_ROLLOUT_PARAMS) {
  extern __shared__ __align__(16) float smem_f[];
  __shared__ double red[(kRolloutBlock / 32) * 14];
  rollout_body<T, POLICY, IDENT, MULTI, Ops>(SN_ROLLOUT_ARGS, smem_f, red);
}

// trees (Ops::kWide): same residency (7 x 64 threads x 144 registers fit the register file), but the full 144
// registers instead of the 128 that ptxas settles on under __launch_bounds__(64, 7): a third of the spills
template <typename T, int POLICY, bool IDENT, bool MULTI, typename Ops>
__global__ void __maxnreg__(144) k_rollout_wide(SN_ROLLOUT_PARAMS) {
  extern __shared__ __align__(16) float smem_f[];
  __shared__ double red[(kRolloutBlock / 32) * 14];
  rollout_body<T, POLICY, IDENT, MULTI, Ops>(SN_ROLLOUT_ARGS, smem_f, red);
}

// ------------------------------------------------------------------------------------------
// Trees of five to eight nodes (supplynet_net.cuh): one lane per node, L lanes per env (L = 5, 6 or 8), 32 / L envs per
// warp (six, five or four), 16 to 24 envs per block.
#define SN_NET_PROLOGUE                                                                                        \
  constexpr int kEnvsPerWarp = 32 / L;                                                                         \
  __shared__ float stage_all[(kNetBlock / 32) * kEnvsPerWarp * 7 * kNF];                                       \
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;                                                  \
  const int64_t warp_env0 = ((int64_t)blockIdx.x * (kNetBlock / 32) + warp) * kEnvsPerWarp;                    \
  const int64_t env = warp_env0 + lane / L;                                                                    \
  const bool valid = env < n && lane < kEnvsPerWarp * L;          /* (spare lanes of a warp: L = 5, 6) */     \
  const int od = 7 * sp.n_obs;                                                                                 \
  float* const stage = stage_all + warp * (kEnvsPerWarp * 7 * kNF);                                            \
  const int n_valid = (int)(n - warp_env0 < kEnvsPerWarp ? (n - warp_env0 > 0 ? n - warp_env0 : 0) : kEnvsPerWarp)

template <typename T, int L>
__global__ void __launch_bounds__(kNetBlock)
k_net_reset(const __grid_constant__ NetSpec<T> sp, int64_t n, int64_t ld, const T* __restrict__ init,
            const T* __restrict__ demand0, T* __restrict__ state, float* __restrict__ obs) {
  SN_NET_PROLOGUE;
  const NetPolicy<T> none{};
  const LaneCfg<T> c = lane_cfg<L>(sp, none, valid);
  LaneEnv<T> e;
  lane_reset(e, c, init, ld, env, demand0);
  lane_store(e, c, state, ld, env);
  if (obs != nullptr) {
    const T cb = group_sum<L>(e.B, c.cust, c.base);
    lane_write_obs<L>(e, c, sp, cb, false, stage, obs + warp_env0 * od, n_valid);
  }
}

template <typename T, int L>
__global__ void __launch_bounds__(kNetBlock)
k_net_observe(const __grid_constant__ NetSpec<T> sp, int64_t n, int64_t ld, const T* __restrict__ state,
              float* __restrict__ obs) {
  SN_NET_PROLOGUE;
  const NetPolicy<T> none{};
  const LaneCfg<T> c = lane_cfg<L>(sp, none, valid);
  LaneEnv<T> e;
  lane_load(e, c, state, ld, env);
  const T cb = group_sum<L>(e.B, c.cust, c.base);
  lane_write_obs<L>(e, c, sp, cb, false, stage, obs + warp_env0 * od, n_valid);
}

// range guard of the integer path for one lane (= one node): see qty_mask in supplynet_core.cuh
template <typename T>
__device__ __forceinline__ uint32_t lane_qty_mask(const LaneEnv<T>& e, T a, T q) {
  if constexpr (std::is_same<T, int32_t>::value)
    return (uint32_t)e.inv | (uint32_t)e.B | (uint32_t)e.unf | (uint32_t)e.dem | (uint32_t)q | (uint32_t)(a ^ (a >> 31));
  else
    return 0u;
}

template <typename T, int L>
__global__ void __launch_bounds__(kNetBlock)
k_net_step(const __grid_constant__ NetSpec<T> sp, int64_t n, int64_t ld, int period_arg, T* __restrict__ state,
           const T* __restrict__ actions, const T* __restrict__ demand, const StepPtrs out) {
  __shared__ double red[(kNetBlock / 32) * 14];
  SN_NET_PROLOGUE;
  const int period = launch_period(out.ctl, period_arg);
  if (period >= sp.t_max) {
    if (out.ctl != nullptr && blockIdx.x == 0 && threadIdx.x == 0) atomicOr(out.ctl + SN_CTL_FLAGS, SN_FLAG_STEP_AFTER_TERMINAL);
    return;
  }
  const bool terminal = period + 1 >= sp.t_max;
  const T* __restrict__ demand_next = out.ctl != nullptr ? demand + (int64_t)(period + 1) * ld * sp.n_src : demand;
  const NetPolicy<T> none{};
  const LaneCfg<T> c = lane_cfg<L>(sp, none, valid);
  LaneEnv<T> e;
  lane_load(e, c, state, ld, env);
  T bj[kNF];
  group_all<L>(e.B, c.base, bj);
  const T a = (c.active && c.slot >= 0) ? __ldg(actions + env * sp.n_agents + c.slot) : T(0);
  T taken;
  const T q = lane_order<SN_POLICY_ACTIONS>(e, c, sp, none, masked_sum<L>(bj, c.cust), a, taken);
  double hold, back;
  lane_step<L>(e, c, sp, q, lane_demand(c, demand_next, ld, env), terminal, bj, hold, back);
  const double reward = group_reward<L>(hold, back, c.base, sp.exact_costs != 0);
  const T cb = masked_sum<L>(bj, c.cust);
  const bool over = c.active && qty_mask_over(lane_qty_mask(e, a, q));
  lane_store(e, c, state, ld, env);
  const bool head = valid && c.k == 0;                 // one lane per env reports
  if (c.active && out.cost) {
    out.cost[(int64_t)c.k * ld + env] = hold;
    out.cost[(int64_t)(kNF + c.k) * ld + env] = back;
  }
  if (head) {
    if (out.reward) out.reward[env] = (float)reward;
    if (out.reward64) out.reward64[env] = reward;
    if (out.ep_return) out.ep_return[env] += reward;
  }
  if (out.obs != nullptr) lane_write_obs<L>(e, c, sp, cb, terminal, stage, out.obs + warp_env0 * od, n_valid);
  if (over && out.ctl != nullptr) atomicOr(out.ctl + SN_CTL_FLAGS, SN_FLAG_QTY_RANGE);
  if (out.stats != nullptr) {
    double v[14];
#pragma unroll
    for (int i = 0; i < 14; ++i) v[i] = 0.0;          // [2..9]: per-facility sums are not kept for these networks
    v[0] = head ? 1.0 : 0.0;
    v[1] = head ? reward : 0.0;
    const bool ep = head && terminal && out.ep_return != nullptr;
    const double ret = ep ? out.ep_return[env] : 0.0;
    v[10] = ep ? 1.0 : 0.0;
    v[11] = ret;
    v[12] = ret * ret;
    v[13] = over ? 1.0 : 0.0;                         // node-steps here (a lane is a node)
    block_reduce_add<14, kNetBlock / 32>(v, out.stats, red);
  }
  if (out.ctl != nullptr) {
    __syncthreads();
    if (threadIdx.x == 0 && last_block_ticket(out.ctl)) {
      out.ctl[SN_CTL_TICKET] = 0;
      out.ctl[SN_CTL_PERIOD] = period + 1;
    }
  }
}

template <typename T, int POLICY, int L>
__global__ void __launch_bounds__(kNetBlock)
k_net_rollout(const __grid_constant__ NetSpec<T> sp, const __grid_constant__ NetPolicy<T> pol, int64_t n, int64_t ld,
              int period0, int n_steps, T* __restrict__ state, const T* __restrict__ demand,
              const T* __restrict__ actions, const RolloutPtrs out, const T* __restrict__ init,
              float* __restrict__ obs0) {
  __shared__ double red[(kNetBlock / 32) * 14];
  SN_NET_PROLOGUE;
  const LaneCfg<T> c = lane_cfg<L>(sp, pol, valid);
  uint32_t qmask = 0;
  const int64_t dstride = ld * sp.n_src, act_stride = n * sp.n_agents;
  const bool head = valid && c.k == 0;
  const bool acts = POLICY == SN_POLICY_ACTIONS && c.active && c.slot >= 0;
  const int64_t act_off = env * sp.n_agents + c.slot;
  LaneEnv<T> e;
  T bj[kNF];                                   // backlog of the eight upstream arcs of this env (same copy in every lane)
  if (init != nullptr) lane_reset(e, c, init, ld, env, demand);
  else lane_load(e, c, state, ld, env);
  group_all<L>(e.B, c.base, bj);
  T cb = masked_sum<L>(bj, c.cust);
  if (init != nullptr && obs0 != nullptr) lane_write_obs<L>(e, c, sp, cb, false, stage, obs0 + warp_env0 * od, n_valid);
  double total = 0.0;
  // inputs of the next period are requested before this one is computed
  T a_pre = acts ? __ldg(actions + act_off) : T(0);
  T d_pre = lane_demand(c, demand + (int64_t)(period0 + 1) * dstride, ld, env);
  for (int s = 0; s < n_steps; ++s) {
    const int t = period0 + s;
    const bool terminal = (t + 1 >= sp.t_max);
    const T a = a_pre, dn = d_pre;
    if (s + 1 < n_steps) {
      if (acts) a_pre = __ldg(actions + (int64_t)(s + 1) * act_stride + act_off);
      d_pre = lane_demand(c, demand + (int64_t)(t + 2) * dstride, ld, env);
    }
    T taken = T(0);
    const T q = lane_order<POLICY>(e, c, sp, pol, cb, a, taken);
    if (out.traj_actions != nullptr && c.active && c.slot >= 0)
      out.traj_actions[(int64_t)s * act_stride + act_off] = (float)taken;
    double hold, back;
    lane_step<L>(e, c, sp, q, dn, terminal, bj, hold, back);
    qmask |= lane_qty_mask(e, a, q);
    cb = weighted_sum<L>(bj, c.cust_m);
    const double r = group_reward<L>(hold, back, c.base, sp.exact_costs != 0);
    total += r;
    if (out.traj_reward != nullptr && head) out.traj_reward[(int64_t)s * n + env] = (float)r;
    if (out.traj_obs != nullptr)
      lane_write_obs<L>(e, c, sp, cb, terminal, stage, out.traj_obs + ((int64_t)s * n + warp_env0) * od, n_valid);
  }
  const bool ended = (period0 + n_steps >= sp.t_max);
  if (out.obs != nullptr) lane_write_obs<L>(e, c, sp, cb, ended, stage, out.obs + warp_env0 * od, n_valid);
  if (state != nullptr) lane_store(e, c, state, ld, env);
  double ep_total = total;
  if (head && out.ep_return) {
    ep_total = (init != nullptr ? 0.0 : out.ep_return[env]) + total;
    out.ep_return[env] = ep_total;
  }
  if (out.stats != nullptr) {
    double v[14];
#pragma unroll
    for (int i = 0; i < 14; ++i) v[i] = 0.0;
    v[0] = head ? (double)n_steps : 0.0;
    v[1] = head ? total : 0.0;
    const bool ep = head && ended && (init != nullptr || out.ep_return != nullptr || period0 == 0);
    v[10] = ep ? 1.0 : 0.0;
    v[11] = ep ? ep_total : 0.0;
    v[12] = ep ? ep_total * ep_total : 0.0;
    v[13] = (c.active && qty_mask_over(qmask)) ? 1.0 : 0.0;      // nodes that left the exact int32 range
    block_reduce_add<14, kNetBlock / 32>(v, out.stats, red);
  }
}
#undef SN_NET_PROLOGUE

// base-stock heuristic for observations of more than four facilities (same arithmetic as k_heuristic, one column at a time)
__global__ void __launch_bounds__(kBlock)
k_heuristic_wide(const __grid_constant__ sn_policy pol, int ncols, int64_t n, const float* __restrict__ obs,
                 float* __restrict__ actions) {
  const int64_t env = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  if (env >= n) return;
  const float* p = obs + (pol.row0_broadcast ? 0 : env * 7 * ncols);
  for (int c = 0; c < ncols; ++c) {
    const float* o = p + 7 * c;
    const float on_order = __fadd_rn(__fadd_rn(__fadd_rn(__ldg(o + 3), __ldg(o + 4)), __ldg(o + 5)), __ldg(o + 6));
    const float ip = __fsub_rn(__fadd_rn(__ldg(o), on_order), __ldg(o + 1));
    double q = pol.target[c] - (double)ip;
    if (pol.rule == SN_RULE_D_PLUS_A) q -= (double)__ldg(o + 2);
    actions[env * ncols + c] = (float)fmin(fmax(q, pol.lb), pol.ub);
  }
}

// BaseStockPolicy.get_order_quantity, ndarray path (utils/heuristics.py:43-57,75), per env.
// float32 sums of the observation terms, float64 subtraction from the target and clip, as numpy
// evaluates `int64_targets - (f32 + f32 - f32) [- f32]`.
__global__ void __launch_bounds__(kBlock)
k_heuristic(const __grid_constant__ sn_policy pol, int ncols, int64_t n, const float* __restrict__ obs,
            float* __restrict__ actions) {
  const int64_t env = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  if (env >= n) return;
  const int od = 7 * ncols;
  const float* p = obs + (pol.row0_broadcast ? 0 : env * od);
  float o[28];
  if (ncols == 4 && !pol.row0_broadcast) {
#pragma unroll
    for (int j = 0; j < 7; ++j) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(p) + j);
      o[4 * j] = v.x; o[4 * j + 1] = v.y; o[4 * j + 2] = v.z; o[4 * j + 3] = v.w;
    }
  } else {
#pragma unroll
    for (int j = 0; j < 28; ++j) o[j] = (j < od) ? __ldg(p + j) : 0.f;
  }
  float r[kF];
#pragma unroll
  for (int c = 0; c < kF; ++c) {
    const float on_order = __fadd_rn(__fadd_rn(__fadd_rn(o[7 * c + 3], o[7 * c + 4]), o[7 * c + 5]), o[7 * c + 6]);
    const float ip = __fsub_rn(__fadd_rn(o[7 * c], on_order), o[7 * c + 1]);
    double q = pol.target[c] - (double)ip;
    if (pol.rule == SN_RULE_D_PLUS_A) q -= (double)o[7 * c + 2];
    q = fmin(fmax(q, pol.lb), pol.ub);
    r[c] = (float)q;
  }
  float* g = actions + env * ncols;
  if (ncols == 4) {
    *reinterpret_cast<float4*>(g) = make_float4(r[0], r[1], r[2], r[3]);
  } else {
#pragma unroll
    for (int c = 0; c < kF; ++c) if (c < ncols) g[c] = r[c];
  }
}

}  // namespace sn

// ==========================================================================================
// Host side: validation, spec lowering, launches.
// ==========================================================================================
namespace sn_detail {
static thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
const char* last_error() { return g_err; }
}  // namespace sn_detail

namespace {
using sn_detail::fail;

// Specialised builds (specialize.py) serve one quantity type and distribution trees of up to four nodes only
#ifndef SN_ONLY_DTYPE
#define SN_ONLY_DTYPE -1
#endif
#if SN_ONLY_DTYPE < 0 || SN_ONLY_DTYPE == 0
#define SN_CASE_I32(EXPR) case SN_I32: return EXPR;
#else
#define SN_CASE_I32(EXPR)
#endif
#if SN_ONLY_DTYPE < 0 || SN_ONLY_DTYPE == 1
#define SN_CASE_F32(EXPR) case SN_F32: return EXPR;
#else
#define SN_CASE_F32(EXPR)
#endif
#if SN_ONLY_DTYPE < 0 || SN_ONLY_DTYPE == 2
#define SN_CASE_F64(EXPR) case SN_F64: return EXPR;
#else
#define SN_CASE_F64(EXPR)
#endif

bool is_integral(double x) { return std::isfinite(x) && std::floor(x) == x && std::fabs(x) < 2147483647.0; }

// the beer game's lead times (envs.py:355-392): tuned kernels; anything else takes the generic path
bool is_tree(const sn_spec* s) { return s->topology == SN_TOPO_TREE; }
// trees of five to eight nodes: shared-memory kernels (supplynet_net.cuh)
bool is_net(const sn_spec* s) { return is_tree(s) && s->n_facilities > SN_MAX_CHAIN; }

bool is_standard(const sn_spec* s) {
  if (s->n_facilities != 4 || is_tree(s)) return false;
  const int li[4] = {2, 2, 2, 1};
  for (int k = 0; k < 4; ++k)
    if (s->info_leadtime[k] != li[k] || s->ship_leadtime[k] != 2) return false;
  return true;
}

template <typename T> sn::DevSpec<T> lower(const sn_spec* s);

int validate(const sn_spec* s, int dtype) {
  if (s == nullptr) return fail(SN_ERR_INVALID, "spec is NULL");
  if (dtype != SN_I32 && dtype != SN_F32 && dtype != SN_F64) return fail(SN_ERR_INVALID, "unknown dtype %d", dtype);
  if (s->n_facilities < 1 || s->n_facilities > (s->topology == SN_TOPO_TREE ? SN_MAX_FACILITIES : SN_MAX_CHAIN))
    return fail(SN_ERR_UNSUPPORTED, "serial chains of 1..4 facilities and trees of 1..8 nodes are supported (n_facilities=%d)", s->n_facilities);
  for (int k = 0; k < s->n_facilities; ++k) {
    const int li = s->info_leadtime[k], ls = s->ship_leadtime[k];
    if (li < 1 || li > 4 || ls < 1 || ls > 4)
      return fail(SN_ERR_UNSUPPORTED, "arc %d: lead times (%d,%d) outside 1..4", k, li, ls);
    if (li + ls > 5)
      return fail(SN_ERR_UNSUPPORTED, "arc %d: information + shipment lead time %d+%d > 5; the reference keeps five pipeline "
                                      "slots (network_components.py:168) and fails its own on_order assertion beyond that", k, li, ls);
  }
  if (is_tree(s) && s->n_items > 1)
    return fail(SN_ERR_UNSUPPORTED, "multi-item batches are supported for serial chains only");
  if (s->n_agents < 1 || s->n_agents > s->n_facilities) return fail(SN_ERR_INVALID, "n_agents=%d out of range", s->n_agents);
  int seen = 0, lo = SN_MAX_FACILITIES, hi = -1;
  for (int i = 0; i < s->n_agents; ++i) {
    const int k = s->agents[i];
    if (k < 0 || k >= s->n_facilities) return fail(SN_ERR_INVALID, "agents[%d]=%d out of range", i, k);
    if (seen & (1 << k)) return fail(SN_ERR_INVALID, "facility %d listed twice", k);
    seen |= 1 << k;
    lo = k < lo ? k : lo;
    hi = k > hi ? k : hi;
  }
  if (s->topology != SN_TOPO_CHAIN && s->topology != SN_TOPO_TREE) return fail(SN_ERR_INVALID, "unknown topology %d", s->topology);
  if (is_tree(s)) {
    const int nf = s->n_facilities;
    int pos_seen = 0, n_src = 0, plo = SN_MAX_FACILITIES, phi = -1;
    for (int k = 0; k < nf; ++k) {
      const int sup = s->supplier[k], op = s->order_pos[k];
      if (sup != SN_EXTERNAL && (sup < 0 || sup >= nf || sup == k)) return fail(SN_ERR_INVALID, "supplier[%d]=%d out of range", k, sup);
      if (op < 0 || op >= nf || (pos_seen & (1 << op))) return fail(SN_ERR_INVALID, "order_pos must be a permutation of 0..%d", nf - 1);
      pos_seen |= 1 << op;
      n_src += s->demand_source[k] ? 1 : 0;
    }
    for (int k = 0; k < nf; ++k) {
      const int sup = s->supplier[k];
      if (sup >= 0 && s->order_pos[sup] <= s->order_pos[k])
        return fail(SN_ERR_INVALID, "node %d must order before its supplier %d (utils/graph.py:5-32)", k, sup);
      int same = 0, lower = 0;                 // ranks among the customers of one supplier: 0..m-1, distinct
      for (int c = 0; c < nf; ++c)
        if (s->supplier[c] == sup) {
          ++same;
          if (c != k && s->customer_rank[c] == s->customer_rank[k]) return fail(SN_ERR_INVALID, "customer_rank of nodes %d and %d collide", k, c);
          lower += s->customer_rank[c] < s->customer_rank[k] ? 1 : 0;
        }
      if (s->customer_rank[k] != lower || s->customer_rank[k] >= same) return fail(SN_ERR_INVALID, "customer_rank[%d]=%d is not a rank among its supplier's customers", k, s->customer_rank[k]);
    }
    if (n_src < 1) return fail(SN_ERR_INVALID, "a tree needs at least one demand source");
    for (int i = 0; i < s->n_agents; ++i) {
      const int op = s->order_pos[s->agents[i]];
      plo = op < plo ? op : plo;
      phi = op > phi ? op : phi;
    }
    if (phi - plo + 1 != s->n_agents)
      return fail(SN_ERR_INVALID, "agent-managed facilities must be contiguous in the order sequence "
                                  "(the reference crashes otherwise, supplynetwork.py:199)");
  } else if (hi - lo + 1 != s->n_agents)
    return fail(SN_ERR_INVALID, "agent-managed facilities must be a contiguous block of the chain "
                                "(the reference crashes otherwise, supplynetwork.py:199)");
  if (s->max_episode_steps < 1) return fail(SN_ERR_INVALID, "max_episode_steps=%d", s->max_episode_steps);
  if (s->rule != SN_RULE_A && s->rule != SN_RULE_D_PLUS_A) return fail(SN_ERR_INVALID, "unknown rule %d", s->rule);
  if (s->n_items < 0 || s->n_items > SN_MAX_ITEMS) return fail(SN_ERR_INVALID, "n_items=%d outside 0..%d", s->n_items, SN_MAX_ITEMS);
  if (dtype == SN_I32) {
    for (int k = 0; k < s->n_facilities; ++k)
      if (!is_integral(s->coplayer_target[k])) return fail(SN_ERR_TYPE, "coplayer_target[%d] must be integral for SN_I32", k);
    if (!is_integral(s->coplayer_lb) || !is_integral(s->coplayer_ub)) return fail(SN_ERR_TYPE, "co-player bounds must be integral for SN_I32");
    if (s->rule == SN_RULE_D_PLUS_A && (!is_integral(s->dpa_offset) || !is_integral(s->dpa_lb) || !is_integral(s->dpa_ub)))
      return fail(SN_ERR_TYPE, "d+a offset/bounds must be integral for SN_I32");
  }
#ifdef SN_STATIC_TOPO
  {   // a specialised build serves exactly the network it was compiled for (specialize.py)
    if (!is_tree(s) || s->n_facilities > sn::kF) return fail(SN_ERR_UNSUPPORTED, "this specialised build serves its own distribution tree only");
    const auto d = lower<int32_t>(s);
    bool same = d.nf == sn::st_nf() && d.max_rank == sn::st_max_rank();
    for (int k = 0; k < sn::kF; ++k)
      same = same && d.sup[k] == sn::st_sup(k) && d.rank[k] == sn::st_rank(k) && d.lat_src[k] == sn::st_lat(k) &&
             d.li[k] == sn::st_li(k) && d.ls[k] == sn::st_ls(k);
    if (!same) return fail(SN_ERR_UNSUPPORTED, "this specialised build was compiled for another network");
  }
#endif
  return SN_OK;
}

template <typename T> T big_quantity();
template <> int32_t big_quantity<int32_t>() { return 1 << 29; }
template <> float big_quantity<float>() { return 1e30f; }
template <> double big_quantity<double>() { return 1e30; }

template <typename T>
sn::DevSpec<T> lower(const sn_spec* s) {
  sn::DevSpec<T> d;
  std::memset(&d, 0, sizeof(d));
  const int nf = s->n_facilities;
  constexpr int KF = sn::kF;                    // 4; 8 in a library specialised for a tree of five to eight nodes
  const int std_li[8] = {2, 2, 2, 1, 1, 1, 1, 1};
  int lo = KF;
  d.nf = nf;
  d.big = big_quantity<T>();
  for (int k = 0; k < KF; ++k) {
    const bool real = k < nf;
    // phantom nodes (k >= nf): a target far below any inventory position makes their order clip to 0,
    // zero cost coefficients keep them out of the reward, any valid lead times will do
    d.target[k] = real ? (T)s->coplayer_target[k] : -d.big;
    d.h[0][k] = real ? s->holding_cost[k] : 0.0;
    d.b[0][k] = real ? s->backorder_cost[k] : 0.0;
    for (int i = 1; i < SN_MAX_ITEMS; ++i) {
      d.h[i][k] = real ? s->item_holding_cost[i - 1][k] : 0.0;
      d.b[i][k] = real ? s->item_backorder_cost[i - 1][k] : 0.0;
    }
    d.li[k] = real ? s->info_leadtime[k] : std_li[k];
    d.ls[k] = real ? s->ship_leadtime[k] : 2;
    d.slot[k] = -1;
    d.pos[k] = -1;
  }
  d.n_items = s->n_items > 1 ? s->n_items : 1;
  for (int i = 0; i < s->n_agents; ++i) {
    d.slot[s->agents[i]] = i;
    lo = s->agents[i] < lo ? s->agents[i] : lo;
  }
  d.clb = (T)s->coplayer_lb; d.cub = (T)s->coplayer_ub;
  d.off = (T)s->dpa_offset; d.lb = (T)s->dpa_lb; d.ub = (T)s->dpa_ub;
  if (s->global_observable) {
    d.n_obs = nf;
    for (int k = 0; k < nf; ++k) d.pos[k] = k;
  } else {
    d.n_obs = s->n_agents;
    for (int i = 0; i < s->n_agents; ++i) d.pos[s->agents[i]] = i;
  }
  d.n_agents = s->n_agents;
  d.lo = lo;
  d.rule = s->rule;
  d.t_max = s->max_episode_steps;
  // topology tables (supplynet_tree.cuh); a chain is the tree k -> k+1 with the retailer as the only source
  const bool tree = is_tree(s);
  int first_agent_pos = KF;
  for (int i = 0; i < s->n_agents; ++i) {
    const int op = tree ? s->order_pos[s->agents[i]] : s->agents[i];
    first_agent_pos = op < first_agent_pos ? op : first_agent_pos;
  }
  d.n_src = 0;
  d.max_rank = 0;
  for (int k = 0; k < KF; ++k) {
    const bool real = k < nf;
    d.sup[k] = !real ? -1 : tree ? s->supplier[k] : (k + 1 < nf ? k + 1 : -1);
    d.rank[k] = (real && tree && d.sup[k] >= 0) ? s->customer_rank[k] : 0;   // the external supplier serves all at once
    if (d.rank[k] > d.max_rank) d.max_rank = d.rank[k];
    const bool source = real && (tree ? s->demand_source[k] != 0 : k == 0);
    d.src[k] = source ? d.n_src++ : -1;
    d.pre[k] = real && (tree ? s->order_pos[k] : k) < first_agent_pos ? 1 : 0;
    d.lat_src[k] = -1;
  }
  for (int k = 0; k < KF; ++k) {
    for (int j = 0; j < 4; ++j) d.m_ls[k][j] = (T)(j == d.ls[k] - 1 ? 1 : 0);
    for (int j = 0; j < 3; ++j) d.m_li[k][j] = (T)(j == d.li[k] - 2 ? 1 : 0);
    d.m_li1[k] = (T)(d.li[k] == 1 ? 1 : 0);
    for (int j = 0; j < KF; ++j) d.m_sup[k][j] = (T)(d.sup[k] == j ? 1 : 0);
    d.m_ext[k] = (T)(d.sup[k] < 0 ? 1 : 0);
    for (int r = 0; r < KF - 1; ++r) d.m_rank[r][k] = (T)(d.rank[k] == r ? 1 : 0);
  }
  for (int j = 0; j < nf; ++j) {          // the customer latest in the order sequence provides latest_demand
    int best = -1;
    for (int c = 0; c < nf; ++c)
      if (d.sup[c] == j && (best < 0 || (tree ? s->order_pos[c] > s->order_pos[best] : c > best))) best = c;
    d.lat_src[j] = best;
  }
  for (int j = 0; j < KF; ++j)
    for (int k = 0; k < KF; ++k) d.m_lat[j][k] = (T)(d.lat_src[j] == k ? 1 : 0);
  return d;
}

// centralized, full layout: action column k = node k and all four nodes observed in chain order
bool is_identity(const sn_spec* s) {
  if (s->n_facilities != 4 || s->n_agents != 4) return false;
  for (int k = 0; k < 4; ++k) if (s->agents[k] != k) return false;
  return true;      // obs: global or local both give positions 0..3
}

template <typename T>
int lower_policy(const sn_policy* p, const sn_spec* s, int dtype, sn::DevPolicy<T>* out) {
  std::memset(out, 0, sizeof(*out));
  if (p == nullptr) return SN_OK;
  for (int i = 0; i < s->n_agents; ++i) {
    if (dtype == SN_I32 && !is_integral(p->target[i])) return fail(SN_ERR_TYPE, "policy target[%d] must be integral for SN_I32", i);
    out->target[s->agents[i]] = (T)p->target[i];       // per-column -> per-node
  }
  if (dtype == SN_I32 && (!is_integral(p->lb) || !is_integral(p->ub))) return fail(SN_ERR_TYPE, "policy bounds must be integral for SN_I32");
  out->lb = (T)p->lb; out->ub = (T)p->ub; out->rule = p->rule;
  return SN_OK;
}

template <typename T>
sn::NetSpec<T> lower_net(const sn_spec* s) {
  sn::NetSpec<T> d;
  std::memset(&d, 0, sizeof(d));
  const int nf = s->n_facilities;
  d.nf = nf;
  d.n_agents = s->n_agents;
  d.rule = s->rule;
  d.t_max = s->max_episode_steps;
  d.big = big_quantity<T>();
  d.clb = (T)s->coplayer_lb; d.cub = (T)s->coplayer_ub;
  d.off = (T)s->dpa_offset; d.lb = (T)s->dpa_lb; d.ub = (T)s->dpa_ub;
  int first_agent_pos = SN_MAX_FACILITIES;
  for (int i = 0; i < s->n_agents; ++i)
    first_agent_pos = s->order_pos[s->agents[i]] < first_agent_pos ? s->order_pos[s->agents[i]] : first_agent_pos;
  for (int k = 0; k < sn::kNF; ++k) {
    const bool real = k < nf;
    d.target[k] = real ? (T)s->coplayer_target[k] : T(0);
    d.h[k] = real ? s->holding_cost[k] : 0.0;
    d.b[k] = real ? s->backorder_cost[k] : 0.0;
    d.li[k] = real ? s->info_leadtime[k] : 1;
    d.ls[k] = real ? s->ship_leadtime[k] : 1;
    d.sup[k] = real ? s->supplier[k] : -1;
    d.slot[k] = d.pos[k] = -1;
    d.src[k] = (real && s->demand_source[k]) ? d.n_src++ : -1;
    d.pre[k] = real && s->order_pos[k] < first_agent_pos ? 1 : 0;
  }
  for (int i = 0; i < s->n_agents; ++i) d.slot[s->agents[i]] = i;
  if (s->global_observable) {
    d.n_obs = nf;
    for (int k = 0; k < nf; ++k) d.pos[k] = k;
  } else {
    d.n_obs = s->n_agents;
    for (int i = 0; i < s->n_agents; ++i) d.pos[s->agents[i]] = i;
  }
  // integer quantities and cost coefficients on a 2^-10 grid: every cost term and partial sum is exact in double
  d.exact_costs = std::is_same<T, int32_t>::value ? 1 : 0;
  for (int k = 0; k < nf; ++k) {
    const double hs = s->holding_cost[k] * 1024.0, bs = s->backorder_cost[k] * 1024.0;
    // |quantity| < 2^31 times |coefficient * 2^10| < 2^18 stays below 2^49: sixteen such terms sum exactly in a double
    if (!(std::fabs(hs) < 262144.0 && std::fabs(bs) < 262144.0 && std::floor(hs) == hs && std::floor(bs) == bs)) d.exact_costs = 0;
  }
  for (int k = 0; k < nf; ++k) {
    int best = -1;
    for (int c = 0; c < nf; ++c) {
      if (s->supplier[c] == k) {                          // c is a customer of k
        d.cust[k] |= 1u << c;
        if (best < 0 || s->order_pos[c] > s->order_pos[best]) best = c;
      }
      if (c != k && s->supplier[k] >= 0 && s->supplier[c] == s->supplier[k] && s->customer_rank[c] < s->customer_rank[k])
        d.before[k] |= 1u << c;                           // served before k by the same supplier
    }
    d.lat_src[k] = best;                                  // the customer latest in the order sequence (supplynetwork.py:193)
  }
  return d;
}

template <typename T>
int lower_net_policy(const sn_policy* p, const sn_spec* s, int dtype, sn::NetPolicy<T>* out) {
  std::memset(out, 0, sizeof(*out));
  if (p == nullptr) return SN_OK;
  for (int i = 0; i < s->n_agents; ++i) {
    if (dtype == SN_I32 && !is_integral(p->target[i])) return fail(SN_ERR_TYPE, "policy target[%d] must be integral for SN_I32", i);
    out->target[s->agents[i]] = (T)p->target[i];
  }
  if (dtype == SN_I32 && (!is_integral(p->lb) || !is_integral(p->ub))) return fail(SN_ERR_TYPE, "policy bounds must be integral for SN_I32");
  out->lb = (T)p->lb; out->ub = (T)p->ub; out->rule = p->rule;
  return SN_OK;
}

int check_batch(int64_t n, int64_t ld) {
  if (n < 1) return fail(SN_ERR_INVALID, "n_envs=%lld must be >= 1", (long long)n);
  if (ld < n || (ld % 4) != 0) return fail(SN_ERR_INVALID, "ld=%lld must be >= n_envs and a multiple of 4", (long long)ld);
  if ((n + sn::kBlock - 1) / sn::kBlock > 2147483647LL) return fail(SN_ERR_INVALID, "n_envs too large");
  return SN_OK;
}

int cuda_ok(const char* what) {
  const cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) return fail(SN_ERR_CUDA, "%s: %s", what, cudaGetErrorString(err));
  return SN_OK;
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
// TMA bulk stores need a 16-byte aligned destination for every warp span (32*od*4 bytes is
// always a multiple of 16) and a 16-byte multiple size.
inline int tma_ok(const float* obs) { return obs != nullptr && aligned16(obs) ? 1 : 0; }

constexpr size_t kObsSmem = sn::kWarpsPerBlock * 2 * 32 * sn::kObsMax * sizeof(float);        // 28,672 B
constexpr size_t kRolloutSmem = (sn::kRolloutBlock / 32) * 2 * 32 * sn::kObsMax * sizeof(float); // 14,336 B

inline unsigned grid_for(int64_t n, int block = sn::kBlock) { return (unsigned)((n + block - 1) / block); }
// eight-node builds stage 56-float observation rows: more than the 48 KB of dynamic shared memory a kernel gets unasked
template <typename K>
inline void allow_smem(K kern, size_t bytes) {
  if (bytes > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}
// net kernels: eight lanes per env
// lanes per env by node count (sn::net_lanes): 5, 6 or 8; a block of four warps holds 4 x (32 / L) envs
inline unsigned net_grid(int64_t n, int lanes) {
  const int64_t per_block = (int64_t)(sn::kNetBlock / 32) * (32 / lanes);
  return (unsigned)((n + per_block - 1) / per_block);
}
#define SN_NET_LANES(NF, CALL)            \
  do {                                    \
    switch (sn::net_lanes(NF)) {          \
      case 5: { CALL(5); } break;         \
      case 6: { CALL(6); } break;         \
      default: { CALL(8); } break;        \
    }                                     \
  } while (0)

// launch with <IDENT, Ops> chosen from the spec: tuned standard kernels (identity specialised or not)
// or the generic lead-time kernels (always the non-identity code path: fewer instantiations)
#ifdef SN_TREE_ONLY
// (specialised builds also compile the centralized full layout - action column k = node k, all four nodes observed in
// node order - with vector action loads and observation stores, like the tuned chain)
#define SN_DISPATCH_TOPO(spec, CALL)                                   \
  do {                                                                 \
    if (!is_tree(spec)) return fail(SN_ERR_UNSUPPORTED, "this specialised build serves its own distribution tree only"); \
    if constexpr (sn::kF == 4) {                                       \
      if (is_identity(spec)) { CALL(true, sn::TreeOps<T>); break; }    \
    }                                                                  \
    CALL(false, sn::TreeOps<T>);                                       \
  } while (0)
#else
#define SN_DISPATCH_TOPO(spec, CALL)                                   \
  do {                                                                 \
    if (is_tree(spec)) { CALL(false, sn::TreeOps<T>); }                \
    else if (!is_standard(spec)) { CALL(false, sn::GenOps<T>); }       \
    else if (is_identity(spec)) { CALL(true, sn::StdOps<T>); }         \
    else { CALL(false, sn::StdOps<T>); }                               \
  } while (0)
#endif

template <typename T>
int do_reset(const sn_spec* spec, int64_t n, int64_t ld, const void* init, const void* demand0, void* state,
             float* obs, cudaStream_t st) {
#ifndef SN_TREE_ONLY
  if (is_net(spec)) {
#define SN_CALL(LN) sn::k_net_reset<T, LN><<<net_grid(n, LN), sn::kNetBlock, 0, st>>>(lower_net<T>(spec), n, ld, (const T*)init, (const T*)demand0, (T*)state, obs)
    SN_NET_LANES(spec->n_facilities, SN_CALL);
#undef SN_CALL
    return cuda_ok("sn_reset launch");
  }
#endif
  const auto d = lower<T>(spec);
#define SN_CALL(ID, OPS) allow_smem(sn::k_reset<T, ID, OPS>, kObsSmem); sn::k_reset<T, ID, OPS><<<grid_for(n), sn::kBlock, kObsSmem, st>>>(d, n, ld, (const T*)init, (const T*)demand0, (T*)state, obs, tma_ok(obs))
  SN_DISPATCH_TOPO(spec, SN_CALL);
#undef SN_CALL
  return cuda_ok("sn_reset launch");
}

template <typename T>
int do_observe(const sn_spec* spec, int64_t n, int64_t ld, const void* state, float* obs, cudaStream_t st) {
#ifndef SN_TREE_ONLY
  if (is_net(spec)) {
#define SN_CALL(LN) sn::k_net_observe<T, LN><<<net_grid(n, LN), sn::kNetBlock, 0, st>>>(lower_net<T>(spec), n, ld, (const T*)state, obs)
    SN_NET_LANES(spec->n_facilities, SN_CALL);
#undef SN_CALL
    return cuda_ok("sn_observe launch");
  }
#endif
  const auto d = lower<T>(spec);
#define SN_CALL(ID, OPS) allow_smem(sn::k_observe<T, ID, OPS>, kObsSmem); sn::k_observe<T, ID, OPS><<<grid_for(n), sn::kBlock, kObsSmem, st>>>(d, n, ld, (const T*)state, obs, tma_ok(obs))
  SN_DISPATCH_TOPO(spec, SN_CALL);
#undef SN_CALL
  return cuda_ok("sn_observe launch");
}

// SMs of the current device (the persistent tiled step launches one CTA per SM)
int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v < 1) v = 148;
    cached[dev] = v;
  }
  return cached[dev];
}

// SN_STEP_KERNEL=plain|tiled forces one implementation of the period step (A/B runs); default: by batch size
int step_kernel_choice() {      // read per call: tests switch it between launches
  const char* v = std::getenv("SN_STEP_KERNEL");
  return (v && !std::strcmp(v, "plain")) ? 1 : (v && !std::strcmp(v, "tiled")) ? 2 : 0;
}
#ifndef SN_TILED_MIN_BYTES
// Batches whose per-period traffic fits the 126 MB L2 are served faster by the plain kernel (shorter prologue, the
// state never leaves L2); beyond that the stream comes from HBM and the tiled kernel wins (measured cross-over:
// 262,144 envs with the observation output, profiles/r2_notes.md)
#define SN_TILED_MIN_BYTES (112ll << 20)
#endif

// cuTensorMapEncodeTiled, fetched from the driver at run time (libcuda is not linked)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}
template <typename T> CUtensorMapDataType tensor_dtype();
template <> CUtensorMapDataType tensor_dtype<int32_t>() { return CU_TENSOR_MAP_DATA_TYPE_INT32; }
template <> CUtensorMapDataType tensor_dtype<float>() { return CU_TENSOR_MAP_DATA_TYPE_FLOAT32; }
template <> CUtensorMapDataType tensor_dtype<double>() { return CU_TENSOR_MAP_DATA_TYPE_FLOAT64; }

template <typename T, bool IDENT, typename Ops, bool ONEPASS = false>
int launch_step_tiled(const sn::DevSpec<T>& d, int64_t n, int64_t ld, int period, void* state, const void* actions,
                      const void* demand, const sn::StepPtrs& out, cudaStream_t st) {
  using C = sn::TileCfg<T, Ops>;
  constexpr int kStages = SN_TILED_STAGES;
  auto kern = &sn::k_step_tiled<T, IDENT, Ops, kStages, ONEPASS>;
  const size_t smem = C::smem_bytes(kStages);
  static bool opted_in[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !opted_in[dev]) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
      return fail(SN_ERR_CUDA, "sn_step: %zu bytes of shared memory per block are not available", smem);
    opted_in[dev] = true;
  }
  // the state matrix [words][ld] as a 2-D tensor: dim0 = env (contiguous), dim1 = state word (stride ld elements);
  // one TMA copy moves a [words][kTile] box
  const EncodeTiledFn enc = encode_tiled();
  if (enc == nullptr) return fail(SN_ERR_CUDA, "sn_step: cuTensorMapEncodeTiled is not available from this driver");
  alignas(64) CUtensorMap tm;
  const cuuint64_t gdim[2] = {(cuuint64_t)ld, (cuuint64_t)C::kWords};
  const cuuint64_t gstride[1] = {(cuuint64_t)ld * sizeof(T)};
  const cuuint32_t box[2] = {(cuuint32_t)C::kTile, (cuuint32_t)C::kWords};
  const cuuint32_t estride[2] = {1, 1};
  const CUresult rc = enc(&tm, tensor_dtype<T>(), 2, state, gdim, gstride, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS) return fail(SN_ERR_CUDA, "sn_step: cuTensorMapEncodeTiled failed (%d)", (int)rc);
  const int64_t tiles = (n + C::kTile - 1) / C::kTile;
  const int64_t sms = sm_count();
  const unsigned grid = (unsigned)(tiles < sms ? tiles : sms);
  if (out.vn_nobs != nullptr) {
    if (tiles > sms || tiles * sn::kVnSlot * 2 > SN_VN_ONEPASS_SCRATCH_DOUBLES) return fail(SN_ERR_UNSUPPORTED, "sn_step_ex: the one-kernel VecNormalize mode takes at most %lld envs on this device (one tile per SM)", (long long)(sms * C::kTile));
    if (!aligned16(out.vn_nobs) != !tma_ok(out.obs) || (out.vn_nrew != nullptr && !aligned16(out.vn_nrew)))
      return fail(SN_ERR_INVALID, "sn_step_ex: obs / vn_nobs / vn_nrew must share their 16-byte alignment");
  }
  static const bool no_coop = [] { const char* v = std::getenv("SN_NO_COOP"); return v && v[0] == '1'; }();   // experiments only
  if (out.vn_nobs != nullptr && out.vn_scratch != nullptr && !no_coop) {
    // blocks meet at a grid barrier: cooperative launch (the runtime guarantees that all of them are resident)
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(C::kThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    const cudaError_t err = cudaLaunchKernelEx(&cfg, kern, d, tm, n, ld, period, (const T*)actions, (const T*)demand, out, tma_ok(out.obs));
    if (err != cudaSuccess) return fail(SN_ERR_CUDA, "sn_step cooperative launch: %s", cudaGetErrorString(err));
    return cuda_ok("sn_step launch");
  }
  kern<<<grid, C::kThreads, smem, st>>>(d, tm, n, ld, period, (const T*)actions, (const T*)demand, out, tma_ok(out.obs));
  return cuda_ok("sn_step launch");
}

template <typename T>
int do_step(const sn_spec* spec, int64_t n, int64_t ld, int period, void* state, const void* actions,
            const void* demand, const sn::StepPtrs& out, cudaStream_t st) {
#ifndef SN_TREE_ONLY
  if (is_net(spec)) {
#define SN_CALL(LN) sn::k_net_step<T, LN><<<net_grid(n, LN), sn::kNetBlock, 0, st>>>(lower_net<T>(spec), n, ld, period, (T*)state, (const T*)actions, (const T*)demand, out)
    SN_NET_LANES(spec->n_facilities, SN_CALL);
#undef SN_CALL
    return cuda_ok("sn_step launch");
  }
#endif
  const auto d = lower<T>(spec);
#ifndef SN_TREE_ONLY
  // the TMA-tiled persistent kernel for HBM-sized batches (and whenever the VecNormalize statistics are fused in: beer-game
  // lead times only), the plain one-thread-per-env kernel for small ones
  const bool tiled_ok = aligned16(state) && aligned16(demand) && aligned16(actions) &&
                        (out.reward == nullptr || aligned16(out.reward)) && (out.env_reward == nullptr || aligned16(out.env_reward)) &&
                        (out.ep_return == nullptr || aligned16(out.ep_return)) && (out.vn_returns == nullptr || aligned16(out.vn_returns));
  const int choice = step_kernel_choice();
  const int words = is_tree(spec) ? sn::kTStateWords : is_standard(spec) ? sn::kStateWords : sn::kGStateWords;
  const int64_t bytes_per_env = 2 * (int64_t)words * (int64_t)sizeof(T) + 24 + (out.obs != nullptr ? 7 * d.n_obs * 4 : 0);
  const bool want_tiled = out.vn_scratch != nullptr || out.vn_nobs != nullptr || choice == 2 || (choice == 0 && n * bytes_per_env >= SN_TILED_MIN_BYTES);
  if ((out.vn_scratch != nullptr || out.vn_nobs != nullptr) && !(tiled_ok && is_standard(spec)))
    return fail(SN_ERR_UNSUPPORTED, "sn_step_ex: fused VecNormalize statistics need the beer game's lead times and 16-byte aligned buffers");
  if (tiled_ok && want_tiled) {
    if (is_tree(spec)) return launch_step_tiled<T, false, sn::TreeOps<T>>(d, n, ld, period, state, actions, demand, out, st);
    if (!is_standard(spec)) return launch_step_tiled<T, false, sn::GenOps<T>>(d, n, ld, period, state, actions, demand, out, st);
    if constexpr (sizeof(T) == 4) {
      if (out.vn_nobs != nullptr) {          // one-kernel VecNormalize: its own instantiation
        if (is_identity(spec)) return launch_step_tiled<T, true, sn::StdOps<T>, true>(d, n, ld, period, state, actions, demand, out, st);
        return launch_step_tiled<T, false, sn::StdOps<T>, true>(d, n, ld, period, state, actions, demand, out, st);
      }
    }
    if (out.vn_nobs != nullptr) return fail(SN_ERR_UNSUPPORTED, "sn_step_ex: the one-kernel VecNormalize mode takes float32 / int32 quantities");
    if (is_identity(spec)) return launch_step_tiled<T, true, sn::StdOps<T>>(d, n, ld, period, state, actions, demand, out, st);
    return launch_step_tiled<T, false, sn::StdOps<T>>(d, n, ld, period, state, actions, demand, out, st);
  }
#endif
#define SN_CALL(ID, OPS) allow_smem(sn::k_step<T, ID, OPS>, kObsSmem); sn::k_step<T, ID, OPS><<<grid_for(n), sn::kBlock, kObsSmem, st>>>(d, n, ld, period, (T*)state, (const T*)actions, (const T*)demand, out, tma_ok(out.obs))
  SN_DISPATCH_TOPO(spec, SN_CALL);
#undef SN_CALL
  return cuda_ok("sn_step launch");
}

template <typename T, int POLICY, bool IDENT, bool MULTI, typename Ops>
void launch_rollout(const sn::DevSpec<T>& d, const sn::DevPolicy<T>& p, int64_t n, int64_t ld, int period0, int n_steps,
                    void* state, const void* demand, const void* actions, const sn::RolloutPtrs& out, int tma,
                    cudaStream_t st, const void* init, float* obs0) {
  auto kern = [] {
    if constexpr (Ops::kWide) return &sn::k_rollout_wide<T, POLICY, IDENT, MULTI, Ops>;
    else return &sn::k_rollout<T, POLICY, IDENT, MULTI, Ops>;
  }();
  if (kRolloutSmem > 48 * 1024) {            // only with experimental block sizes (SN_ROLLOUT_BLOCK > 192)
    static bool opted_in = false;
    if (!opted_in) {
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kRolloutSmem);
      opted_in = true;
    }
  }
  kern<<<grid_for(n, sn::kRolloutBlock), sn::kRolloutBlock, kRolloutSmem, st>>>(
      d, p, n, ld, period0, n_steps, (T*)state, (const T*)demand, (const T*)actions, out, tma, (const T*)init, obs0);
}

template <typename T, int POLICY>
void launch_rollout_for(const sn_spec* spec, const sn::DevSpec<T>& d, const sn::DevPolicy<T>& p, int64_t n, int64_t ld,
                        int period0, int n_steps, void* state, const void* demand, const void* actions,
                        const sn::RolloutPtrs& out, int tma, cudaStream_t st, const void* init, float* obs0) {
#define SN_ARGS d, p, n, ld, period0, n_steps, state, demand, actions, out, tma, st, init, obs0
#ifdef SN_TREE_ONLY
  if constexpr (sn::kF == 4) {
    if (is_tree(spec) && is_identity(spec)) { launch_rollout<T, POLICY, true, false, sn::TreeOps<T>>(SN_ARGS); return; }
  }
#endif
  if (is_tree(spec)) launch_rollout<T, POLICY, false, false, sn::TreeOps<T>>(SN_ARGS);
#ifndef SN_TREE_ONLY
  else if (!is_standard(spec)) {
    if (d.n_items > 1) launch_rollout<T, POLICY, false, true, sn::GenOps<T>>(SN_ARGS);
    else launch_rollout<T, POLICY, false, false, sn::GenOps<T>>(SN_ARGS);
  }
  else if (is_identity(spec)) {
    if (d.n_items > 1) launch_rollout<T, POLICY, true, true, sn::StdOps<T>>(SN_ARGS);
    else launch_rollout<T, POLICY, true, false, sn::StdOps<T>>(SN_ARGS);
  } else {
    if (d.n_items > 1) launch_rollout<T, POLICY, false, true, sn::StdOps<T>>(SN_ARGS);
    else launch_rollout<T, POLICY, false, false, sn::StdOps<T>>(SN_ARGS);
  }
#endif
#undef SN_ARGS
}

template <typename T>
int do_rollout(const sn_spec* spec, int dtype, int64_t n, int64_t ld, int period0, int n_steps, void* state,
               const void* demand, int policy_kind, const void* actions, const sn_policy* policy,
               const sn::RolloutPtrs& out, cudaStream_t st, const void* init = nullptr, float* obs0 = nullptr) {
#ifndef SN_TREE_ONLY
  if (is_net(spec)) {
    sn::NetPolicy<T> np;
    const int nrc = lower_net_policy<T>(policy, spec, dtype, &np);
    if (nrc != SN_OK) return nrc;
    const auto nd = lower_net<T>(spec);
    if (policy_kind == SN_POLICY_ACTIONS) {
#define SN_CALL(LN) sn::k_net_rollout<T, SN_POLICY_ACTIONS, LN><<<net_grid(n, LN), sn::kNetBlock, 0, st>>>( \
          nd, np, n, ld, period0, n_steps, (T*)state, (const T*)demand, (const T*)actions, out, (const T*)init, obs0)
      SN_NET_LANES(spec->n_facilities, SN_CALL);
#undef SN_CALL
    } else {
#define SN_CALL(LN) sn::k_net_rollout<T, SN_POLICY_BASE_STOCK, LN><<<net_grid(n, LN), sn::kNetBlock, 0, st>>>( \
          nd, np, n, ld, period0, n_steps, (T*)state, (const T*)demand, nullptr, out, (const T*)init, obs0)
      SN_NET_LANES(spec->n_facilities, SN_CALL);
#undef SN_CALL
    }
    return cuda_ok("sn_rollout launch");
  }
#endif
  const auto d = lower<T>(spec);
  sn::DevPolicy<T> p;
  const int rc = lower_policy<T>(policy, spec, dtype, &p);
  if (rc != SN_OK) return rc;
  const int tma = (out.traj_obs == nullptr || aligned16(out.traj_obs)) && (out.obs == nullptr || aligned16(out.obs)) &&
                  (obs0 == nullptr || aligned16(obs0)) && (((int64_t)n * 7 * d.n_obs * 4) % 16 == 0);
  if (policy_kind == SN_POLICY_ACTIONS)
    launch_rollout_for<T, SN_POLICY_ACTIONS>(spec, d, p, n, ld, period0, n_steps, state, demand, actions, out, tma, st, init, obs0);
  else
    launch_rollout_for<T, SN_POLICY_BASE_STOCK>(spec, d, p, n, ld, period0, n_steps, state, demand, nullptr, out, tma, st, init, obs0);
  return cuda_ok("sn_rollout launch");
}

}  // namespace

extern "C" {

int32_t sn_abi_version(void) { return SN_ABI_VERSION; }

const char* sn_last_error(void) { return sn_detail::last_error(); }

int32_t sn_obs_dim(const sn_spec* spec) {
  if (spec == nullptr) return fail(SN_ERR_INVALID, "spec is NULL");
  if (spec->n_facilities < 1 || spec->n_facilities > (is_tree(spec) ? SN_MAX_FACILITIES : SN_MAX_CHAIN))
    return fail(SN_ERR_UNSUPPORTED, "n_facilities=%d", spec->n_facilities);
  if (spec->n_agents < 1 || spec->n_agents > spec->n_facilities) return fail(SN_ERR_INVALID, "n_agents=%d out of range", spec->n_agents);
  return SN_OBS_PER_FACILITY * (spec->global_observable ? spec->n_facilities : spec->n_agents);
}

int32_t sn_spec_validate(const sn_spec* spec, int32_t dtype) { return validate(spec, dtype); }

int32_t sn_state_words(const sn_spec* spec) {
  if (spec == nullptr) return fail(SN_ERR_INVALID, "spec is NULL");
  return is_net(spec) ? SN_NET_STATE_WORDS : is_tree(spec) ? SN_TREE_STATE_WORDS : is_standard(spec) ? SN_STATE_WORDS : SN_GEN_STATE_WORDS;
}

int32_t sn_init_words(const sn_spec* spec) {
  if (spec == nullptr) return fail(SN_ERR_INVALID, "spec is NULL");
  return is_net(spec) ? SN_NET_INIT_WORDS : is_standard(spec) ? SN_INIT_WORDS : SN_GEN_INIT_WORDS;
}

int32_t sn_reset(const sn_spec* spec, int32_t dtype, int64_t n_envs, int64_t ld, const void* init,
                 const void* demand0, void* state, float* obs, void* stream) {
  int rc = validate(spec, dtype);
  if (rc != SN_OK) return rc;
  if ((rc = check_batch(n_envs, ld)) != SN_OK) return rc;
  if (!init || !demand0 || !state) return fail(SN_ERR_INVALID, "sn_reset: init, demand0 and state must be non-NULL");
  cudaStream_t st = (cudaStream_t)stream;
  switch (dtype) {
    SN_CASE_I32(do_reset<int32_t>(spec, n_envs, ld, init, demand0, state, obs, st))
    SN_CASE_F32(do_reset<float>(spec, n_envs, ld, init, demand0, state, obs, st))
    SN_CASE_F64(do_reset<double>(spec, n_envs, ld, init, demand0, state, obs, st))
  }
  return fail(SN_ERR_UNSUPPORTED, "this build of the library does not serve dtype %d", dtype);
}

int32_t sn_observe(const sn_spec* spec, int32_t dtype, int64_t n_envs, int64_t ld, const void* state, float* obs,
                   void* stream) {
  int rc = validate(spec, dtype);
  if (rc != SN_OK) return rc;
  if ((rc = check_batch(n_envs, ld)) != SN_OK) return rc;
  if (!state || !obs) return fail(SN_ERR_INVALID, "sn_observe: state and obs must be non-NULL");
  cudaStream_t st = (cudaStream_t)stream;
  switch (dtype) {
    SN_CASE_I32(do_observe<int32_t>(spec, n_envs, ld, state, obs, st))
    SN_CASE_F32(do_observe<float>(spec, n_envs, ld, state, obs, st))
    SN_CASE_F64(do_observe<double>(spec, n_envs, ld, state, obs, st))
  }
  return fail(SN_ERR_UNSUPPORTED, "this build of the library does not serve dtype %d", dtype);
}

int32_t sn_step_ex(const sn_spec* spec, int32_t dtype, int64_t n_envs, int64_t ld, const sn_step_io* io, void* stream) {
  int rc = validate(spec, dtype);
  if (rc != SN_OK) return rc;
  if ((rc = check_batch(n_envs, ld)) != SN_OK) return rc;
  if (io == nullptr) return fail(SN_ERR_INVALID, "sn_step_ex: io is NULL");
  if (io->ctl == nullptr) {
    if (io->period < 0) return fail(SN_ERR_INVALID, "period=%d", io->period);
    if (io->period >= spec->max_episode_steps)
      return fail(SN_ERR_TERMINAL, "Cannot take action when the state is terminal.");     // envs.py:88-89
  }
  if (!io->state || !io->actions || !io->demand) return fail(SN_ERR_INVALID, "sn_step: state, actions and demand must be non-NULL");
  if (spec->n_agents == 4 && !aligned16(io->actions)) return fail(SN_ERR_INVALID, "sn_step: actions must be 16-byte aligned");
  if (io->env_reward != nullptr) {
    if (is_tree(spec) || (spec->n_items != 2 && spec->n_items != 4))
      return fail(SN_ERR_UNSUPPORTED, "sn_step_ex: env_reward needs a multi-item chain batch with 2 or 4 items (n_items=%d)", spec->n_items);
    if (n_envs % spec->n_items != 0) return fail(SN_ERR_INVALID, "sn_step_ex: n_envs=%lld is not a multiple of n_items=%d", (long long)n_envs, spec->n_items);
  }
  const bool vn = io->vn_scratch || io->vn_obs_rms || io->vn_ret_rms || io->vn_returns || io->vn_nobs || io->vn_nrew;
  if (vn) {
    if (!io->vn_obs_rms) return fail(SN_ERR_INVALID, "sn_step_ex: the VecNormalize extras need vn_obs_rms");
    if (!io->vn_scratch && !io->vn_nobs) return fail(SN_ERR_INVALID, "sn_step_ex: give vn_scratch (update), vn_nobs (normalise) or both");
    if (io->vn_returns && !io->vn_ret_rms) return fail(SN_ERR_INVALID, "sn_step_ex: vn_returns needs vn_ret_rms");
    if (io->vn_nrew && (!io->vn_nobs || !io->vn_ret_rms)) return fail(SN_ERR_INVALID, "sn_step_ex: vn_nrew needs vn_nobs and vn_ret_rms");
    if (io->vn_returns && !io->vn_scratch) return fail(SN_ERR_INVALID, "sn_step_ex: vn_returns needs vn_scratch (the statistics update)");
    if (io->vn_nobs && dtype == SN_F64) return fail(SN_ERR_UNSUPPORTED, "sn_step_ex: the one-kernel VecNormalize mode takes SN_I32 / SN_F32 quantities");
    if (!io->ctl || !io->obs) return fail(SN_ERR_INVALID, "sn_step_ex: fused VecNormalize statistics need ctl and obs");
    if (io->vn_returns && !io->reward) return fail(SN_ERR_INVALID, "sn_step_ex: vn_returns needs reward");
    if (spec->n_items > 1 || !is_standard(spec)) return fail(SN_ERR_UNSUPPORTED, "sn_step_ex: fused VecNormalize statistics need the single-item beer-game chain");
  }
  sn::StepPtrs out{};
  out.obs = io->obs; out.reward = io->reward; out.reward64 = io->reward64; out.cost = io->cost;
  out.ep_return = io->ep_return; out.stats = io->stats; out.env_reward = io->env_reward; out.ctl = io->ctl;
  out.vn_scratch = io->vn_scratch; out.vn_obs_rms = io->vn_obs_rms; out.vn_ret_rms = io->vn_ret_rms;
  out.vn_returns = io->vn_returns; out.vn_gamma = io->vn_gamma;
  out.vn_nobs = io->vn_nobs; out.vn_nrew = io->vn_nrew; out.vn_clip_obs = io->vn_clip_obs; out.vn_clip_rew = io->vn_clip_reward;
  out.vn_eps = io->vn_epsilon;
  cudaStream_t st = (cudaStream_t)stream;
  switch (dtype) {
    SN_CASE_I32(do_step<int32_t>(spec, n_envs, ld, io->period, io->state, io->actions, io->demand, out, st))
    SN_CASE_F32(do_step<float>(spec, n_envs, ld, io->period, io->state, io->actions, io->demand, out, st))
    SN_CASE_F64(do_step<double>(spec, n_envs, ld, io->period, io->state, io->actions, io->demand, out, st))
  }
  return fail(SN_ERR_UNSUPPORTED, "this build of the library does not serve dtype %d", dtype);
}

int32_t sn_step(const sn_spec* spec, int32_t dtype, int64_t n_envs, int64_t ld, int32_t period, void* state,
                const void* actions, const void* demand_next, float* obs, float* reward, double* reward64,
                double* cost, double* ep_return, double* stats, void* stream) {
  sn_step_io io;
  std::memset(&io, 0, sizeof(io));
  io.state = state; io.actions = actions; io.demand = demand_next; io.period = period;
  io.obs = obs; io.reward = reward; io.reward64 = reward64; io.cost = cost; io.ep_return = ep_return; io.stats = stats;
  return sn_step_ex(spec, dtype, n_envs, ld, &io, stream);
}

int32_t sn_heuristic(const sn_policy* policy, int32_t n_cols, int64_t n_envs, const float* obs, float* actions,
                     void* stream) {
  if (!policy || !obs || !actions) return fail(SN_ERR_INVALID, "sn_heuristic: NULL argument");
  if (n_cols < 1 || n_cols > SN_MAX_FACILITIES) return fail(SN_ERR_INVALID, "n_cols=%d out of range", n_cols);
  if (n_envs < 1) return fail(SN_ERR_INVALID, "n_envs=%lld must be >= 1", (long long)n_envs);
  if (policy->rule != SN_RULE_A && policy->rule != SN_RULE_D_PLUS_A) return fail(SN_ERR_INVALID, "unknown rule %d", policy->rule);
  if (n_cols == 4 && (!aligned16(obs) || !aligned16(actions))) return fail(SN_ERR_INVALID, "sn_heuristic: obs/actions must be 16-byte aligned");
  if (n_cols > SN_MAX_CHAIN) sn::k_heuristic_wide<<<grid_for(n_envs), sn::kBlock, 0, (cudaStream_t)stream>>>(*policy, n_cols, n_envs, obs, actions);
  else sn::k_heuristic<<<grid_for(n_envs), sn::kBlock, 0, (cudaStream_t)stream>>>(*policy, n_cols, n_envs, obs, actions);
  return cuda_ok("sn_heuristic launch");
}

int32_t sn_rollout(const sn_spec* spec, int32_t dtype, int64_t n_envs, int64_t ld, int32_t period0, int32_t n_steps,
                   void* state, const void* demand, int32_t policy_kind, const void* actions,
                   const sn_policy* policy, float* traj_obs, float* traj_reward, float* traj_actions, float* obs,
                   double* ep_return, double* stats, void* stream) {
  int rc = validate(spec, dtype);
  if (rc != SN_OK) return rc;
  if ((rc = check_batch(n_envs, ld)) != SN_OK) return rc;
  if (period0 < 0 || n_steps < 1) return fail(SN_ERR_INVALID, "period0=%d n_steps=%d", period0, n_steps);
  if (period0 >= spec->max_episode_steps)
    return fail(SN_ERR_TERMINAL, "Cannot take action when the state is terminal.");
  if (period0 + n_steps > spec->max_episode_steps)
    return fail(SN_ERR_INVALID, "rollout of %d steps from period %d crosses the episode end (%d)", n_steps, period0, spec->max_episode_steps);
  if (!state || !demand) return fail(SN_ERR_INVALID, "sn_rollout: state and demand must be non-NULL");
  if (policy_kind == SN_POLICY_ACTIONS) {
    if (!actions) return fail(SN_ERR_INVALID, "sn_rollout: actions must be non-NULL for SN_POLICY_ACTIONS");
    if (spec->n_agents == 4 && !aligned16(actions)) return fail(SN_ERR_INVALID, "sn_rollout: actions must be 16-byte aligned");
  } else if (policy_kind == SN_POLICY_BASE_STOCK) {
    if (!policy) return fail(SN_ERR_INVALID, "sn_rollout: policy must be non-NULL for SN_POLICY_BASE_STOCK");
    if (policy->rule != SN_RULE_A && policy->rule != SN_RULE_D_PLUS_A) return fail(SN_ERR_INVALID, "unknown policy rule %d", policy->rule);
  } else {
    return fail(SN_ERR_INVALID, "unknown policy_kind %d", policy_kind);
  }
  if (traj_actions && spec->n_agents == 4 && !aligned16(traj_actions)) return fail(SN_ERR_INVALID, "sn_rollout: traj_actions must be 16-byte aligned");
  const sn::RolloutPtrs out{traj_obs, traj_reward, traj_actions, obs, ep_return, stats};
  cudaStream_t st = (cudaStream_t)stream;
  switch (dtype) {
    SN_CASE_I32(do_rollout<int32_t>(spec, dtype, n_envs, ld, period0, n_steps, state, demand, policy_kind, actions, policy, out, st))
    SN_CASE_F32(do_rollout<float>(spec, dtype, n_envs, ld, period0, n_steps, state, demand, policy_kind, actions, policy, out, st))
    SN_CASE_F64(do_rollout<double>(spec, dtype, n_envs, ld, period0, n_steps, state, demand, policy_kind, actions, policy, out, st))
  }
  return fail(SN_ERR_UNSUPPORTED, "this build of the library does not serve dtype %d", dtype);
}

int32_t sn_episode(const sn_spec* spec, int32_t dtype, int64_t n_envs, int64_t ld, int32_t n_steps, const void* init,
                   const void* demand, int32_t policy_kind, const void* actions, const sn_policy* policy, float* obs0,
                   float* traj_obs, float* traj_reward, float* traj_actions, float* obs, void* state,
                   double* ep_return, double* stats, void* stream) {
  int rc = validate(spec, dtype);
  if (rc != SN_OK) return rc;
  if ((rc = check_batch(n_envs, ld)) != SN_OK) return rc;
  if (n_steps < 1 || n_steps > spec->max_episode_steps)
    return fail(SN_ERR_INVALID, "sn_episode: n_steps=%d outside 1..%d", n_steps, spec->max_episode_steps);
  if (!init || !demand) return fail(SN_ERR_INVALID, "sn_episode: init and demand must be non-NULL");
  if (policy_kind == SN_POLICY_ACTIONS) {
    if (!actions) return fail(SN_ERR_INVALID, "sn_episode: actions must be non-NULL for SN_POLICY_ACTIONS");
    if (spec->n_agents == 4 && !aligned16(actions)) return fail(SN_ERR_INVALID, "sn_episode: actions must be 16-byte aligned");
  } else if (policy_kind == SN_POLICY_BASE_STOCK) {
    if (!policy) return fail(SN_ERR_INVALID, "sn_episode: policy must be non-NULL for SN_POLICY_BASE_STOCK");
    if (policy->rule != SN_RULE_A && policy->rule != SN_RULE_D_PLUS_A) return fail(SN_ERR_INVALID, "unknown policy rule %d", policy->rule);
  } else {
    return fail(SN_ERR_INVALID, "unknown policy_kind %d", policy_kind);
  }
  if (traj_actions && spec->n_agents == 4 && !aligned16(traj_actions)) return fail(SN_ERR_INVALID, "sn_episode: traj_actions must be 16-byte aligned");
  const sn::RolloutPtrs out{traj_obs, traj_reward, traj_actions, obs, ep_return, stats};
  cudaStream_t st = (cudaStream_t)stream;
  switch (dtype) {
    SN_CASE_I32(do_rollout<int32_t>(spec, dtype, n_envs, ld, 0, n_steps, state, demand, policy_kind, actions, policy, out, st, init, obs0))
    SN_CASE_F32(do_rollout<float>(spec, dtype, n_envs, ld, 0, n_steps, state, demand, policy_kind, actions, policy, out, st, init, obs0))
    SN_CASE_F64(do_rollout<double>(spec, dtype, n_envs, ld, 0, n_steps, state, demand, policy_kind, actions, policy, out, st, init, obs0))
  }
  return fail(SN_ERR_UNSUPPORTED, "this build of the library does not serve dtype %d", dtype);
}

}  // extern "C"
