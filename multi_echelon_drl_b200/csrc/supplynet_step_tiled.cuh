// supplynet_step_tiled.cuh — the single-period step (envs.py:82-113 for all envs) as a persistent, TMA-fed
// pipeline.  Included by supplynet_kernels.cu inside namespace sn (after StepPtrs / k_step).
//
// Why: the plain k_step issues 40 scalar loads per thread behind 164-176 registers and runs at 17 % occupancy:
// not enough bytes in flight for the 345-457 B that one env-step moves (profiles/r1_ncu_k_step_summary.json).
// Here one CTA per SM walks over tiles of kTile consecutive envs.  A tile is, in the structure-of-arrays state,
// a [words][kTile] box of the [words][ld] matrix; a producer warp moves it with ONE tensor-map TMA copy
// (cp.async.bulk.tensor.2d, SASS UTMALDG / UTMASTG) into one of kStages shared-memory stages, completion counted on
// an mbarrier, next to the tile's action rows ([kTile][n_agents], contiguous: one cp.async.bulk, UBLKCP) and
// demand row(s).  (A first version moved the box as `words` separate 512-byte bulk copies: the TMA unit then
// spends its time on per-copy overhead, ~84 copies per tile = 3.2 us, 0.31-0.40 of the HBM peak; profiles/r2_notes.md.)  The consumer warps (one thread
// per env) read their column from shared memory, run the same per-env code as every other kernel
// (env_orders / Ops::step), write the new state back IN PLACE, the observation rows ([kTile][obs_dim] f32, one
// contiguous span) and the rewards next to it, and hand the stage back; the producer stores it with bulk
// copies and refills it kStages tiles later.  Loads of kStages-1 tiles are in flight while one is computed, with
// no register cost, and the stores drain asynchronously.
//
// A partial last tile keeps the tensor path for the state box (columns beyond `ld` are clipped by the TMA unit:
// zero-filled on load, dropped on store) and uses per-thread accesses for the row-major tensors.
#pragma once

#ifndef SN_TILED_STAGES
#define SN_TILED_STAGES 4
#endif

// ---- mbarrier + bulk-copy PTX ------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
// bounded: a protocol error traps (the launch fails) instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  for (uint32_t spins = 0; !mbar_try_wait(bar, parity); ++spins)
    if (spins > (1u << 22)) __trap();
}
__device__ __forceinline__ void bulk_load_g2s(void* sdst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(sdst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// [rows][kTile] box of the state matrix at column c0 (tensor map: dim0 = env, dim1 = state word)
__device__ __forceinline__ void tensor_load_2d(void* sdst, const CUtensorMap* tm, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(smem_u32(sdst)), "l"(tm), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tensor_store_2d(const CUtensorMap* tm, int c0, int c1, const void* ssrc) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];"
               ::"l"(tm), "r"(c0), "r"(c1), "r"(smem_u32(ssrc)) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// Low-latency slot entries (one-kernel VecNormalize): a double travels as two 8-byte words, each = 32 bits of the value
// + the 32-bit tag of the launch that wrote it.  8-byte accesses are single-copy atomic, so a word whose tag matches
// carries that launch's half of the value: no fence and no separate flag between writer and reader.
__device__ __forceinline__ void ll_store(double* entry, double v, uint32_t tag) {
  const unsigned long long bits = (unsigned long long)__double_as_longlong(v), t = (unsigned long long)tag << 32;
  const unsigned long long w0 = (bits & 0xffffffffull) | t, w1 = (bits >> 32) | t;
  asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(entry), "l"(w0), "l"(w1) : "memory");
}
__device__ __forceinline__ void ll_load(const double* entry, unsigned long long& w0, unsigned long long& w1) {
  asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(entry) : "memory");
}
__device__ __forceinline__ bool ll_value(unsigned long long w0, unsigned long long w1, uint32_t tag, double& v) {
  v = __longlong_as_double((long long)((w0 & 0xffffffffull) | (w1 << 32)));
  return (uint32_t)(w0 >> 32) == tag && (uint32_t)(w1 >> 32) == tag;
}

constexpr int kVnSlot = 64;          // 16-byte entries per block in the scratch of the one-kernel VecNormalize mode (1 KB)

// ---- stage geometry ----------------------------------------------------------------------------------------
template <typename T, typename Ops>
struct TileCfg {
  static constexpr int kTile = sizeof(T) == 8 ? 64 : 128;          // envs per tile: 512-byte row segments
  static constexpr int kConsumerWarps = kTile / 32;
  static constexpr int kThreads = kTile + 32;                      // + one producer warp
  static constexpr int kWords = Ops::kWords;
  static constexpr int kStateBytes = kWords * kTile * (int)sizeof(T);
  static constexpr int kActBytes = kTile * kF * (int)sizeof(T);
  static constexpr int kDemBytes = Ops::kMaxSrc * kTile * (int)sizeof(T);
  static constexpr int kObsBytes = kTile * 28 * 4;
  static constexpr int kRewBytes = kTile * 4;
  static constexpr int kOffAct = kStateBytes;
  static constexpr int kOffDem = kOffAct + kActBytes;
  static constexpr int kOffObs = kOffDem + kDemBytes;
  static constexpr int kOffRew = kOffObs + kObsBytes;
  static constexpr int kOffERew = kOffRew + kRewBytes;
  static constexpr int kOffEpr = kOffERew + kRewBytes;            // episode returns of the tile (double, in/out)
  static constexpr int kOffVnr = kOffEpr + kTile * 8;              // VecNormalize's discounted returns (double, in/out)
  static constexpr int kStageBytes = kOffVnr + kTile * 8;          // every part a multiple of 128 bytes
  static constexpr int kBarBytes = 128;                            // full[kStages], done[kStages]
  static constexpr int kRedDoubles = (kConsumerWarps + 1) * 14;   // block_reduce_add scratch
  static constexpr int kVnDoubles = 2 * kTile;                      // column-moment partials of the fused VecNormalize
  static constexpr int kStatFloats = 80;                           // one-kernel VecNormalize: mean[32], 1/std[32], 1/std of the returns
  static constexpr int kRedBytes = ((kRedDoubles + kVnDoubles) * 8 + 16 + kStatFloats * 4 + 127) & ~127;
  static constexpr size_t smem_bytes(int stages) { return 128 + kBarBytes + kRedBytes + (size_t)stages * kStageBytes; }
};

// RunningMeanStd.update for one column / the returns: parallel-variance merge of (mean, var, count) with the
// batch moments (sum, sum of squares over bn samples); SURVEY Appendix D, same formulas as vecnorm_kernels.cu.
__device__ __forceinline__ void rms_merge_inplace(double* mean, double* var, double count, double sum, double sumsq, double bn) {
  const double bmean = sum / bn;
  const double bvar = fmax(sumsq / bn - bmean * bmean, 0.0);
  const double delta = bmean - *mean;
  const double tot = count + bn;
  const double m2 = *var * count + bvar * bn + delta * delta * count * bn / tot;
  *mean = *mean + delta * bn / tot;
  *var = m2 / tot;
}

// ONEPASS: the instantiation that carries the one-kernel VecNormalize section (its sixteen loads in flight cost 60
// registers; kept out of the streaming instantiation, whose allocation it would disturb).
template <typename T, bool IDENT, typename Ops, int kStages, bool ONEPASS>
__global__ void __launch_bounds__(TileCfg<T, Ops>::kThreads, 1)
k_step_tiled(const __grid_constant__ DevSpec<T> sp, const __grid_constant__ CUtensorMap tm_state, int64_t n, int64_t ld,
             int period_arg, const T* __restrict__ actions, const T* __restrict__ demand, const StepPtrs out, int obs_tma) {
  using C = TileCfg<T, Ops>;
  constexpr int kTile = C::kTile;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);              // [kStages] producer -> consumers (tx bytes)
  uint64_t* done = full + kStages;                                 // [kStages] consumers -> producer
  double* red = reinterpret_cast<double*>(smem + C::kBarBytes);    // block reductions
  double* vn_red = red + C::kRedDoubles;                           // [2][kTile]
  int* last_flag = reinterpret_cast<int*>(vn_red + C::kVnDoubles);
  uint8_t* stages = smem + C::kBarBytes + C::kRedBytes;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int period = launch_period(out.ctl, period_arg);
  if (period >= sp.t_max) {
    if (out.ctl != nullptr && blockIdx.x == 0 && tid == 0) atomicOr(out.ctl + SN_CTL_FLAGS, SN_FLAG_STEP_AFTER_TERMINAL);
    return;
  }
  const bool terminal = period + 1 >= sp.t_max;
  // one-kernel VecNormalize: statistics (optionally updated across the grid) and normalisation in this launch
  const bool vn_onepass = ONEPASS && out.vn_nobs != nullptr;
  const bool vn_barrier = vn_onepass && out.vn_scratch != nullptr;
  const int seq = vn_barrier ? *reinterpret_cast<const volatile int32_t*>(out.ctl + SN_CTL_SEQ) : 0;
  const int64_t dstride = Ops::dstride(sp, ld);
  const T* __restrict__ demand_next = out.ctl != nullptr ? demand + (int64_t)(period + 1) * dstride : demand;
  const int od = 7 * sp.n_obs, na = sp.n_agents, n_src = Ops::n_src(sp);
  const int64_t n_tiles = (n + kTile - 1) / kTile;
  const int64_t my_tiles = blockIdx.x < n_tiles ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&done[s], C::kConsumerWarps);
    }
    mbar_fence_init();
  }
  fence_proxy_async_smem();
  __syncthreads();

  double acc[14];
#pragma unroll
  for (int i = 0; i < 14; ++i) acc[i] = 0.0;
  double vsx = 0.0, vsxx = 0.0, vsr = 0.0, vsrr = 0.0;             // fused VecNormalize partial sums

  if (warp == C::kConsumerWarps) {
    // ================= producer warp: every lane moves its share of the rows of a tile =====================
    for (int64_t it = 0; it < my_tiles + kStages - 1; ++it) {
      const int64_t j = it - (kStages - 1);                        // tile whose results leave now
      if (j >= 0) {
        const int s = (int)(j % kStages);
        const int64_t env0 = (blockIdx.x + j * gridDim.x) * kTile;
        const bool full_tile = env0 + kTile <= n;
        uint8_t* st = stages + (size_t)s * C::kStageBytes;
        mbar_wait(&done[s], (uint32_t)((j / kStages) & 1));
        if (lane == 0) tensor_store_2d(&tm_state, (int)env0, 0, st);
        if (full_tile) {
          if (lane == 31 && obs_tma) bulk_store_s2g(out.obs + env0 * od, st + C::kOffObs, (uint32_t)(kTile * od * 4));
          if (lane == 30 && out.reward != nullptr) bulk_store_s2g(out.reward + env0, st + C::kOffRew, (uint32_t)(kTile * 4));
          if (lane == 29 && out.env_reward != nullptr)
            bulk_store_s2g(out.env_reward + (env0 >> (sp.n_items == 2 ? 1 : 2)), st + C::kOffERew, (uint32_t)(kTile / sp.n_items * 4));
          if (lane == 28 && out.ep_return != nullptr) bulk_store_s2g(out.ep_return + env0, st + C::kOffEpr, (uint32_t)(kTile * 8));
          if (lane == 27 && out.vn_returns != nullptr) bulk_store_s2g(out.vn_returns + env0, st + C::kOffVnr, (uint32_t)(kTile * 8));
        }
        bulk_commit();
        if (vn_onepass) {
          // the normalised tile is built in the next stage's buffers, after the grid barrier: the step's own results
          // above are already on their way while the blocks meet
          const int s1 = (s + 1) % kStages;
          uint8_t* st1 = stages + (size_t)s1 * C::kStageBytes;
          mbar_wait(&done[s1], (uint32_t)((j / kStages) & 1));
          if (full_tile) {
            if (lane == 26 && obs_tma) bulk_store_s2g(out.vn_nobs + env0 * od, st1 + C::kOffObs, (uint32_t)(kTile * od * 4));
            if (lane == 25 && out.vn_nrew != nullptr) bulk_store_s2g(out.vn_nrew + env0, st1 + C::kOffRew, (uint32_t)(kTile * 4));
          }
          bulk_commit();
        }
      }
      if (it < my_tiles) {
        bulk_wait_read<1>();                                       // the stores of tile it-kStages have left this stage
        __syncwarp();
        const int s = (int)(it % kStages);
        const int64_t env0 = (blockIdx.x + it * gridDim.x) * kTile;
        const int cols = (int)((ld - env0) < kTile ? (ld - env0) : kTile);
        const bool full_tile = env0 + kTile <= n;
        uint8_t* st = stages + (size_t)s * C::kStageBytes;
        const uint32_t row_bytes = (uint32_t)(cols * sizeof(T));
        if (lane == 0) {                                           // the whole box counts, clipped columns included
          uint32_t tx = (uint32_t)C::kStateBytes + (uint32_t)n_src * row_bytes;
          if (full_tile) {
            tx += (uint32_t)(kTile * na * sizeof(T));
            if (out.ep_return != nullptr) tx += (uint32_t)(kTile * 8);
            if (out.vn_returns != nullptr) tx += (uint32_t)(kTile * 8);
          }
          mbar_arrive_expect_tx(&full[s], tx);
          tensor_load_2d(st, &tm_state, (int)env0, 0, &full[s]);
        }
        __syncwarp();
        if (lane == 31 && full_tile) bulk_load_g2s(st + C::kOffAct, actions + env0 * na, (uint32_t)(kTile * na * sizeof(T)), &full[s]);
        // per-env accumulators that live in global memory travel with the tile: a load issued by a consumer warp
        // would stall it for the full memory latency (one warp per scheduler: nothing else to run meanwhile)
        if (lane == 23 && full_tile && out.ep_return != nullptr) bulk_load_g2s(st + C::kOffEpr, out.ep_return + env0, (uint32_t)(kTile * 8), &full[s]);
        if (lane == 22 && full_tile && out.vn_returns != nullptr) bulk_load_g2s(st + C::kOffVnr, out.vn_returns + env0, (uint32_t)(kTile * 8), &full[s]);
        if (lane >= 24 && lane < 24 + n_src)
          bulk_load_g2s(st + C::kOffDem + (size_t)(lane - 24) * kTile * sizeof(T), demand_next + (int64_t)(lane - 24) * ld + env0,
                        row_bytes, &full[s]);
      }
    }
    bulk_wait_read<0>();
  } else {
    // ================= consumer warps: thread = env of the tile ==========================================
    for (int64_t it = 0; it < my_tiles; ++it) {
      const int s = (int)(it % kStages);
      const int64_t env0 = (blockIdx.x + it * gridDim.x) * kTile;
      const int64_t env = env0 + tid;
      const bool valid = env < n, full_tile = env0 + kTile <= n;
      uint8_t* st = stages + (size_t)s * C::kStageBytes;
      T* s_state = reinterpret_cast<T*>(st);
      const T* s_act = reinterpret_cast<const T*>(st + C::kOffAct);
      const T* s_dem = reinterpret_cast<const T*>(st + C::kOffDem);
      float* s_obs = reinterpret_cast<float*>(st + C::kOffObs);
      float* s_rew = reinterpret_cast<float*>(st + C::kOffRew);
      float* s_erew = reinterpret_cast<float*>(st + C::kOffERew);
      double* s_epr = reinterpret_cast<double*>(st + C::kOffEpr);
      double* s_vnr = reinterpret_cast<double*>(st + C::kOffVnr);
      // one-kernel VecNormalize: the running statistics as they are before this step, requested before the tile
      // arrives (read before the grid barrier in any case: block 0 rewrites them after it)
      double vn_m = 0.0, vn_v = 1.0, vn_cnt = 0.0;
      if (vn_onepass) {
        if (tid < od) { vn_m = out.vn_obs_rms[tid]; vn_v = out.vn_obs_rms[od + tid]; vn_cnt = out.vn_obs_rms[2 * od]; }
        if (tid == 32 && out.vn_ret_rms != nullptr) { vn_m = out.vn_ret_rms[0]; vn_v = out.vn_ret_rms[1]; vn_cnt = out.vn_ret_rms[2]; }
      }
      mbar_wait(&full[s], (uint32_t)((it / kStages) & 1));

      typename Ops::E e;
      StepOut so;
      so.reward = 0.0;
      bool over = false;
#pragma unroll
      for (int k = 0; k < kF; ++k) so.hold[k] = so.back[k] = 0.0;
      if (valid) {
        Ops::load(sp, s_state, kTile, tid, e);
        T a[kF] = {T(0), T(0), T(0), T(0)};
        if (full_tile) load_actions_smem<IDENT>(s_act, tid, sp, a);
        else load_actions<IDENT>(actions, env, sp, a);
        const typename Ops::D dn = Ops::load_demand_smem(s_dem, kTile, tid, sp);
        T q[kF];
        env_orders<IDENT>(e, sp, a, q);
        const uint32_t om = qty_mask_orders(a, q);
        if (sp.n_items > 1) Ops::template step<true>(e, sp, chain_item(sp, env), q, dn, terminal, so);
        else Ops::template step<false>(e, sp, 0, q, dn, terminal, so);
        over = qty_mask_over(om | qty_mask_state<T>(e));
        Ops::store(s_state, kTile, tid, e);
        const float rf = (float)so.reward;
        if (out.reward != nullptr) { if (full_tile) s_rew[tid] = rf; else out.reward[env] = rf; }
        if (out.reward64) out.reward64[env] = so.reward;
        if (out.cost) {
#pragma unroll
          for (int k = 0; k < kF; ++k) {
            out.cost[k * ld + env] = so.hold[k];
            out.cost[(kF + k) * ld + env] = so.back[k];
          }
        }
        double ret_ep = 0.0;
        if (out.ep_return) {
          ret_ep = (full_tile ? s_epr[tid] : out.ep_return[env]) + so.reward;
          if (full_tile) s_epr[tid] = ret_ep; else out.ep_return[env] = ret_ep;
        }
        if (out.vn_returns != nullptr) {                           // VecNormalize: ret <- ret*gamma + r (the float reward)
          const double r = (full_tile ? s_vnr[tid] : out.vn_returns[env]) * out.vn_gamma + (double)rf;
          if (full_tile) s_vnr[tid] = r; else out.vn_returns[env] = r;
          vsr += r;
          vsrr += r * r;
        }
        if (out.obs != nullptr) emit_obs<IDENT>(e, sp, terminal, s_obs + tid * od);
        acc[0] += 1.0;
        acc[1] += so.reward;
#pragma unroll
        for (int k = 0; k < 4; ++k) { acc[2 + k] += so.hold[k]; acc[6 + k] += so.back[k]; }
        if (terminal && out.ep_return != nullptr) { acc[10] += 1.0; acc[11] += ret_ep; acc[12] += ret_ep * ret_ep; }
        if (over) acc[13] += 1.0;
      }
      if (out.env_reward != nullptr) {                             // every lane takes part in the shuffles
        const double tot = env_reward_sum(so.reward, sp.n_items);
        const int sh = sp.n_items == 2 ? 1 : 2;                    // n_items is 2 or 4 here
        if (valid && (tid & (sp.n_items - 1)) == 0) {
          if (full_tile) s_erew[tid >> sh] = (float)tot;
          else out.env_reward[env >> sh] = (float)tot;
        }
      }
      const bool obs_by_threads = out.obs != nullptr && !(full_tile && obs_tma);
      if (vn_onepass) {
        // ---- one-kernel VecNormalize (one tile per CTA): moments -> grid barrier -> merged statistics -> normalise
        uint8_t* st1 = stages + (size_t)((s + 1) % kStages) * C::kStageBytes;
        float* s_nobs = reinterpret_cast<float*>(st1 + C::kOffObs);
        float* s_nrew = reinterpret_cast<float*>(st1 + C::kOffRew);
        float* s_stat = reinterpret_cast<float*>(last_flag + 4);                           // mean[32], 1/std[32], 1/std of the returns
        const int rows = (int)((n - env0) < kTile ? (n - env0) : kTile);
        const int ng = od <= 32 ? kTile / od : 0;
        const int ret_thread = 32;                                 // od <= 28 on this kernel: thread 32 owns the return statistics
        // the step's own results can leave now: hand stage s to the producer (it only reads it; so do we from here on)
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&done[s]);
        named_bar_sync(1, kTile);                                  // the tile's observation rows are complete
        double m = vn_m, v = vn_v;
        const double cnt = vn_cnt;
        if (vn_barrier) {
          // Every block publishes the moments of its tile in a slot of its own and reads everybody else's: flag and
          // data travel together (each 8-byte half of an entry carries 32 bits of the value and this launch's tag,
          // the low-latency protocol of collective libraries), so a reader that sees the tag has the value - no
          // fence, no separate flag, ONE trip through the L2 between "my sums are ready" and "I have all sums".
          // Every block adds the slots up in block order (deterministic, unlike atomics).
          double* slots = out.vn_scratch;                          // [gridDim.x][kVnSlot] entries of 16 bytes
          double* mine = slots + (size_t)blockIdx.x * kVnSlot * 2;
          const uint32_t tag = (uint32_t)seq + 1u;
          const bool has_ret = out.vn_returns != nullptr;
          if (tid < ng * od) {
            const int c = tid % od, g = tid / od;
            double sx = 0.0, sxx = 0.0;
#pragma unroll 8
            for (int r = g; r < rows; r += ng) {
              const double x = (double)s_obs[r * od + c];
              sx += x;
              sxx += x * x;
            }
            vn_red[tid] = sx;
            vn_red[kTile + tid] = sxx;
          }
          double sr = vsr, srr = vsrr;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) { sr += __shfl_xor_sync(0xffffffffu, sr, o); srr += __shfl_xor_sync(0xffffffffu, srr, o); }
          if (lane == 0) { red[2 * warp] = sr; red[2 * warp + 1] = srr; }
          named_bar_sync(1, kTile);
          if (tid < od) {
            double a = 0.0, b = 0.0;
            for (int g = 0; g < ng; ++g) { a += vn_red[g * od + tid]; b += vn_red[kTile + g * od + tid]; }
            ll_store(mine + 2 * tid, a, tag);
            ll_store(mine + 2 * (od + tid), b, tag);
          } else if (tid == ret_thread && has_ret) {
            double a = 0.0, b = 0.0;
            for (int w = 0; w < C::kConsumerWarps; ++w) { a += red[2 * w]; b += red[2 * w + 1]; }
            ll_store(mine + 2 * (2 * od), a, tag);
            ll_store(mine + 2 * (2 * od + 1), b, tag);
          }
          named_bar_sync(1, kTile);                                // vn_red is free again (the group sums have been read)
          const double bn = (double)n;
          {
            // all consumer threads fetch: thread = (block group g, column c; c == od: the returns).  Eight blocks per
            // round: sixteen independent loads in flight, then the tags are checked; a round is repeated until every
            // entry carries this launch's tag (bounded: a protocol error traps, never hangs)
            const int c = tid & 31, g = tid >> 5;
            double a = 0.0, b = 0.0;
            if (c < od || (c == od && has_ret)) {
              const int ia = c < od ? c : 2 * od, ib = c < od ? od + c : 2 * od + 1;
              for (int b0 = g; b0 < (int)gridDim.x; b0 += C::kConsumerWarps * 8) {
                double va[8], vb[8];
                bool ok;
                uint32_t spins = 0;
                do {
                  unsigned long long w[8][4];
#pragma unroll
                  for (int u = 0; u < 8; ++u) {
                    const int bq = b0 + C::kConsumerWarps * u;
                    if (bq < (int)gridDim.x) {
                      ll_load(slots + ((size_t)bq * kVnSlot + ia) * 2, w[u][0], w[u][1]);
                      ll_load(slots + ((size_t)bq * kVnSlot + ib) * 2, w[u][2], w[u][3]);
                    }
                  }
                  ok = true;
#pragma unroll
                  for (int u = 0; u < 8; ++u) {
                    const int bq = b0 + C::kConsumerWarps * u;
                    if (bq < (int)gridDim.x) {
                      ok = ok && ll_value(w[u][0], w[u][1], tag, va[u]) && ll_value(w[u][2], w[u][3], tag, vb[u]);
                    }
                  }
                  if (++spins > (1u << 22)) __trap();
                } while (!ok);
#pragma unroll
                for (int u = 0; u < 8; ++u)
                  if (b0 + C::kConsumerWarps * u < (int)gridDim.x) { a += va[u]; b += vb[u]; }
              }
            }
            vn_red[tid] = a;
            vn_red[kTile + tid] = b;
          }
          named_bar_sync(1, kTile);
          if (tid < od || (tid == ret_thread && out.vn_returns != nullptr)) {
            const int c = tid < od ? tid : od;
            double a = 0.0, b = 0.0;
#pragma unroll
            for (int g = 0; g < C::kConsumerWarps; ++g) { a += vn_red[g * 32 + c]; b += vn_red[kTile + g * 32 + c]; }   // fixed order
            rms_merge_inplace(&m, &v, cnt, a, b, bn);
            if (blockIdx.x == 0) {
              if (tid < od) {
                out.vn_obs_rms[tid] = m;
                out.vn_obs_rms[od + tid] = v;
                if (tid == 0) out.vn_obs_rms[2 * od] = cnt + bn;
              } else {
                out.vn_ret_rms[0] = m; out.vn_ret_rms[1] = v; out.vn_ret_rms[2] = cnt + bn;
              }
            }
          }
          if (blockIdx.x == 0 && tid == 0) {                       // everybody has read them: advance period and sequence
            out.ctl[SN_CTL_PERIOD] = period + 1;
            out.ctl[SN_CTL_SEQ] = seq + 1;
          }
        }
        if (tid < od) { s_stat[tid] = (float)m; s_stat[32 + tid] = (float)(1.0 / sqrt(v + out.vn_eps)); }
        if (tid == ret_thread) s_stat[64] = (float)(1.0 / sqrt(v + out.vn_eps));
        named_bar_sync(1, kTile);
        if (valid) {                                               // same arithmetic as sn_vecnorm_apply
          const float co = out.vn_clip_obs;
          for (int j = 0; j < od; j += 4) {                        // od is a multiple of 7; rows are 16-byte aligned when it is one of 4 too
            if ((od & 3) == 0) {
              const float4 x = *reinterpret_cast<const float4*>(s_obs + tid * od + j);
              float4 z;
              z.x = fminf(fmaxf((x.x - s_stat[j]) * s_stat[32 + j], -co), co);
              z.y = fminf(fmaxf((x.y - s_stat[j + 1]) * s_stat[32 + j + 1], -co), co);
              z.z = fminf(fmaxf((x.z - s_stat[j + 2]) * s_stat[32 + j + 2], -co), co);
              z.w = fminf(fmaxf((x.w - s_stat[j + 3]) * s_stat[32 + j + 3], -co), co);
              *reinterpret_cast<float4*>(s_nobs + tid * od + j) = z;
            } else {
              for (int jj = j; jj < j + 4 && jj < od; ++jj)
                s_nobs[tid * od + jj] = fminf(fmaxf((s_obs[tid * od + jj] - s_stat[jj]) * s_stat[32 + jj], -co), co);
            }
          }
          if (out.vn_nrew != nullptr) {
            const float z = fminf(fmaxf((float)so.reward * s_stat[64], -out.vn_clip_rew), out.vn_clip_rew);
            if (full_tile) s_nrew[tid] = z; else out.vn_nrew[env] = z;
          }
        }
        if (obs_by_threads) {
          named_bar_sync(1, kTile);
          for (int i = tid; i < rows * od; i += kTile) { out.obs[env0 * od + i] = s_obs[i]; out.vn_nobs[env0 * od + i] = s_nobs[i]; }
        }
      } else if (obs_by_threads || out.vn_scratch != nullptr) {
        named_bar_sync(1, kTile);                                  // the tile's observation rows are complete
        const int rows = (int)((n - env0) < kTile ? (n - env0) : kTile);
        if (obs_by_threads)                                        // partial tile / unaligned output: coalesced copy
          for (int i = tid; i < rows * od; i += kTile) out.obs[env0 * od + i] = s_obs[i];
        if (out.vn_scratch != nullptr && tid < (od <= 32 ? (kTile / od) * od : 0)) {
          // column moments of this tile: thread = (row group g, column c); groups interleave the rows
          const int c = tid % od, g = tid / od, ng = kTile / od;
#pragma unroll 8
          for (int r = g; r < rows; r += ng) {
            const double x = (double)s_obs[r * od + c];
            vsx += x;
            vsxx += x * x;
          }
        }
      }
      fence_proxy_async_smem();                                    // generic-proxy writes -> visible to the bulk stores
      __syncwarp();
      if (lane == 0) mbar_arrive(&done[vn_onepass ? (s + 1) % kStages : s]);
    }
  }

  // ---- epilogue: statistics, VecNormalize moments, device period counter ---------------------------------
  if (out.stats != nullptr) block_reduce_add<14, C::kConsumerWarps + 1>(acc, out.stats, red);
  if (out.vn_scratch != nullptr && !vn_onepass) {
    // per-thread partials -> one atomic per block and column into scratch; the last block merges them into the
    // running statistics (in place: nothing reads them before the next kernel) and clears scratch
    __syncthreads();
    const int ng = od <= 32 ? kTile / od : 0;
    if (tid < ng * od) { vn_red[tid] = vsx; vn_red[kTile + tid] = vsxx; }
    double sr = vsr, srr = vsrr;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { sr += __shfl_xor_sync(0xffffffffu, sr, o); srr += __shfl_xor_sync(0xffffffffu, srr, o); }
    if (lane == 0) { red[2 * warp] = sr; red[2 * warp + 1] = srr; }
    __syncthreads();
    if (tid < od) {
      double a = 0.0, b = 0.0;
      for (int g = 0; g < ng; ++g) { a += vn_red[g * od + tid]; b += vn_red[kTile + g * od + tid]; }
      atomicAdd(&out.vn_scratch[tid], a);
      atomicAdd(&out.vn_scratch[od + tid], b);
    } else if (tid == 32 + (od & ~31) && out.vn_returns != nullptr) {   // a thread outside the column range
      double a = 0.0, b = 0.0;
      for (int w = 0; w < C::kConsumerWarps; ++w) { a += red[2 * w]; b += red[2 * w + 1]; }
      atomicAdd(&out.vn_scratch[2 * od], a);
      atomicAdd(&out.vn_scratch[2 * od + 1], b);
    }
  }
  if (out.ctl != nullptr && !vn_barrier) {                       // (the grid barrier has advanced the period already)
    __syncthreads();
    if (tid == 0) *last_flag = last_block_ticket(out.ctl) ? 1 : 0;
    __syncthreads();
    if (*last_flag) {
      if (out.vn_scratch != nullptr) {
        __threadfence();
        volatile double* sc = out.vn_scratch;
        const double bn = (double)n;
        if (tid < od) {
          const double count = out.vn_obs_rms[2 * od];
          rms_merge_inplace(&out.vn_obs_rms[tid], &out.vn_obs_rms[od + tid], count, sc[tid], sc[od + tid], bn);
        }
        if (tid == 64 && out.vn_returns != nullptr) {
          rms_merge_inplace(&out.vn_ret_rms[0], &out.vn_ret_rms[1], out.vn_ret_rms[2], sc[2 * od], sc[2 * od + 1], bn);
          out.vn_ret_rms[2] += bn;
        }
        __syncthreads();                                           // every column has read the old count
        if (tid == 0) out.vn_obs_rms[2 * od] += bn;
        if (tid < 2 * od + 2) out.vn_scratch[tid] = 0.0;
      }
      if (tid == 0) {
        out.ctl[SN_CTL_TICKET] = 0;
        out.ctl[SN_CTL_PERIOD] = period + 1;
      }
    }
  }
}
