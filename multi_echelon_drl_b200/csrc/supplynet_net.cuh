// supplynet_net.cuh — distribution trees of FIVE TO EIGHT internal nodes (continued).
//
// Same semantics as supplynet_tree.cuh (every node buys from one supplier, may serve several customers and / or its
// own demand; what SupplyNetwork does for such node / arc lists, supplynetwork.py:182-268), different residency:
// 8 nodes x 16 state words do not fit the register file at one-wave occupancy, so each env lives in SHARED MEMORY
// for the duration of a launch (word-major, thread-minor: `sm[word][thread]`, conflict-free) and the topology is
// applied by plain indexing (`inv[sup[k]]`) with loops over the run-time node count - no mask arithmetic, compact
// code.  One thread per env, 64-thread blocks, (18 x nf x 64) quantity words of dynamic shared memory per block.
//
// State word map in HBM ([SN_NET_STATE_WORDS = 128][ld]; node k owns words 16k .. 16k+15):
//   +0 inv   +1 unf_ind   +2 latI (latest internal demand)   +3..6 U[0..3]   +7..9 info[0..2]   +10..13 ship[0..3]
//   +14 B (backlog of the node's upstream arc)   +15 dem (current external demand)
// Init words ([SN_NET_INIT_WORDS = 72][ld]): inv[8]; so[k][j] at 8+4k+j; sh[k][j] at 40+4k+j (unused = 0).
// Demand tables: one row per demand source and period, as for the smaller trees.
#pragma once
#include "supplynet_core.cuh"

namespace sn {

constexpr int kNF = 8;
constexpr int kNetStateWords = 16 * kNF;     // 128
constexpr int kNetInitWords = kNF + 8 * kNF; // 72
constexpr int kNetBlock = 64;
constexpr int kNetSmemWords = 18;            // per node: 16 state words + this period's order + the fifth pipeline slot

enum { N_INV = 0, N_UNF = 1, N_LATI = 2, N_U = 3, N_INFO = 7, N_SHIP = 10, N_B = 14, N_DEM = 15, N_Q = 16, N_U4 = 17 };

template <typename T>
struct NetSpec {
  int nf, n_agents, n_obs, n_src, rule, t_max;
  T target[kNF];
  T clb, cub, off, lb, ub, big;
  double h[kNF], b[kNF];
  int li[kNF], ls[kNF];
  int sup[kNF];                    // supplier node, -1 = external supplier
  int slot[kNF], pos[kNF];         // action column / observation position of node k, -1 = none
  int src[kNF];                    // demand-table row of node k, -1 = no external demand
  int pre[kNF];                    // orders inside before_action (ahead of the first agent in the order sequence)
  int lat_src[kNF];                // customer whose arriving order is node k's latest demand, -1 = none
  int fill_seq[kNF];               // arcs in serving order: by customer rank (suppliers are independent of each other)
  int cust_start[kNF + 1];         // customers of node k: cust_list[cust_start[k] .. cust_start[k+1])
  int cust_list[kNF];
};

template <typename T>
struct NetPolicy {
  T target[kNF];                   // per NODE
  T lb, ub;
  int rule;
};

// one env in shared memory
template <typename T>
struct NetEnv {
  T* sm;
  __device__ __forceinline__ T& at(int node, int f) const { return sm[(node * kNetSmemWords + f) * kNetBlock + threadIdx.x]; }

  __device__ __forceinline__ T unfilled(const NetSpec<T>& sp, int k) const {       // supplynetwork.py:130-136
    T cb = T(0);
    for (int i = sp.cust_start[k]; i < sp.cust_start[k + 1]; ++i) cb = cb + at(sp.cust_list[i], N_B);
    return cb + (at(k, N_DEM) + at(k, N_UNF));
  }
  __device__ __forceinline__ T latest(int k) const { return at(k, N_LATI) + at(k, N_DEM); }   // :139
  __device__ __forceinline__ T base_stock_order(const NetSpec<T>& sp, int k, T target, T lb, T ub, int rule) const {
    const T U[4] = {at(k, N_U), at(k, N_U + 1), at(k, N_U + 2), at(k, N_U + 3)};
    return base_stock(target, at(k, N_INV), U, unfilled(sp, k), latest(k), lb, ub, rule);
  }
};

template <typename T>
__device__ __forceinline__ void net_load(const NetEnv<T>& e, const NetSpec<T>& sp, const T* __restrict__ state, int64_t ld, int64_t env) {
  for (int k = 0; k < sp.nf; ++k)
#pragma unroll
    for (int f = 0; f < 16; ++f) e.at(k, f) = state[(int64_t)(16 * k + f) * ld + env];
}

template <typename T>
__device__ __forceinline__ void net_store(const NetEnv<T>& e, const NetSpec<T>& sp, T* __restrict__ state, int64_t ld, int64_t env) {
  for (int k = 0; k < sp.nf; ++k)
#pragma unroll
    for (int f = 0; f < 16; ++f) state[(int64_t)(16 * k + f) * ld + env] = e.at(k, f);
}

// current external demand of every node from one period's rows of the demand table
template <typename T>
__device__ __forceinline__ void net_set_demand(const NetEnv<T>& e, const NetSpec<T>& sp, const T* __restrict__ row, int64_t ld, int64_t env) {
  for (int k = 0; k < sp.nf; ++k) e.at(k, N_DEM) = sp.src[k] >= 0 ? __ldg(row + (int64_t)sp.src[k] * ld + env) : T(0);
}

// Arc.reset / Node.reset + before_action(0) (network_components.py:138-168,293-319; supplynetwork.py:182-199)
template <typename T>
__device__ __forceinline__ void net_reset(const NetEnv<T>& e, const NetSpec<T>& sp, const T* __restrict__ init, int64_t ld,
                                          int64_t env, const T* __restrict__ demand0) {
  for (int k = 0; k < sp.nf; ++k) {
    const int li = sp.li[k], ls = sp.ls[k];
    T so[4], sh[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      so[j] = j < li ? __ldg(init + (int64_t)(kNF + 4 * k + j) * ld + env) : T(0);
      sh[j] = j < ls ? __ldg(init + (int64_t)(5 * kNF + 4 * k + j) * ld + env) : T(0);
    }
    e.at(k, N_INV) = __ldg(init + (int64_t)k * ld + env);
    // ([0]*4 + shipments + sales_orders)[::-1][:5]: newest slip first, then shipments latest first
    T seq[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      T v = T(0);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (li - 1 - i == j) v = so[j];
        if (i >= li && ls - 1 - (i - li) == j) v = sh[j];
      }
      seq[i] = v;
    }
    e.at(k, N_U) = seq[0]; e.at(k, N_U + 1) = seq[1]; e.at(k, N_U + 2) = seq[2]; e.at(k, N_U + 3) = seq[3] + seq[4];
    e.at(k, N_B) = so[0];                                    // the slip with remaining lead time 1 arrives now
    e.at(k, N_INFO) = so[1]; e.at(k, N_INFO + 1) = so[2]; e.at(k, N_INFO + 2) = so[3];
#pragma unroll
    for (int j = 0; j < 4; ++j) e.at(k, N_SHIP + j) = sh[j];
    e.at(k, N_UNF) = T(0);
    e.at(k, N_LATI) = T(0);
  }
  for (int k = 0; k < sp.nf; ++k)
    if (sp.lat_src[k] >= 0) e.at(k, N_LATI) = e.at(sp.lat_src[k], N_B);
  net_set_demand(e, sp, demand0, ld, env);
}

// orders of all nodes for this period, into slot N_Q (depend on start-of-period quantities only)
template <int POLICY, typename T>
__device__ __forceinline__ void net_orders(const NetEnv<T>& e, const NetSpec<T>& sp, const NetPolicy<T>& pol,
                                           const T* __restrict__ act_row /* this env's action row or NULL */,
                                           float* __restrict__ act_out /* this env's row of traj_actions or NULL */) {
  for (int k = 0; k < sp.nf; ++k) {
    T x;
    if (sp.slot[k] >= 0) {
      if (POLICY == SN_POLICY_BASE_STOCK) x = e.base_stock_order(sp, k, pol.target[k], pol.lb, pol.ub, pol.rule);
      else x = __ldg(act_row + sp.slot[k]);
      if (act_out != nullptr) act_out[sp.slot[k]] = (float)x;
      if (sp.rule == 1) x = tclamp(e.latest(k) + x + sp.off, sp.lb, sp.ub);          // utils/wrappers.py:25-27
    } else {
      x = e.base_stock_order(sp, k, sp.target[k], sp.clb, sp.cub, 0);                 // network_components.py:361-366
    }
    e.at(k, N_Q) = tmax(x, T(0));                                                     // :368-374
  }
}

// One period (agent_action + after_action + get_cost + before_action unless terminal).  `cost` (may be NULL) receives
// holding[0..7] / backorder[8..15] of this env at stride ld.
template <typename T>
__device__ __forceinline__ double net_step(const NetEnv<T>& e, const NetSpec<T>& sp, const T* __restrict__ demand_next,
                                           int64_t ld, int64_t env, bool terminal, double* __restrict__ cost) {
  // receipts on every arc (supplynetwork.py:243-255)
  for (int k = 0; k < sp.nf; ++k) {
    T u[5] = {e.at(k, N_Q), e.at(k, N_U), e.at(k, N_U + 1), e.at(k, N_U + 2), e.at(k, N_U + 3)};
    const T arr = e.at(k, N_SHIP);
    e.at(k, N_INV) = e.at(k, N_INV) + arr;
    T rem = arr;
#pragma unroll
    for (int i = 4; i >= 0; --i) {
      const T f = tmin(u[i], rem);
      u[i] = u[i] - f;
      rem = rem - f;
    }
    e.at(k, N_SHIP) = e.at(k, N_SHIP + 1); e.at(k, N_SHIP + 1) = e.at(k, N_SHIP + 2);
    e.at(k, N_SHIP + 2) = e.at(k, N_SHIP + 3); e.at(k, N_SHIP + 3) = T(0);
    e.at(k, N_U) = u[0]; e.at(k, N_U + 1) = u[1]; e.at(k, N_U + 2) = u[2]; e.at(k, N_U + 3) = u[3];
    e.at(k, N_U4) = u[4];
  }
  // fills: every supplier serves its customer arcs in rank order from what it has (Arc.fill_orders, :260-262)
  for (int i = 0; i < sp.nf; ++i) {
    const int k = sp.fill_seq[i], s = sp.sup[k];
    const T avail = s >= 0 ? e.at(s, N_INV) : sp.big;
    const T amt = tmin(avail, e.at(k, N_B));
    if (s >= 0) e.at(s, N_INV) = avail - amt;
    e.at(k, N_B) = e.at(k, N_B) - amt;
    e.at(k, N_SHIP + sp.ls[k] - 1) = e.at(k, N_SHIP + sp.ls[k] - 1) + amt;
  }
  // independent demand (network_components.py:323-345) and cost (supplynetwork.py:285-310)
  double ch = 0.0, cb = 0.0;
  for (int k = 0; k < sp.nf; ++k) {
    if (sp.src[k] >= 0) {
      const T tot = e.at(k, N_DEM) + e.at(k, N_UNF);
      const T sold = tmax(T(0), tmin(e.at(k, N_INV), tot));
      e.at(k, N_INV) = e.at(k, N_INV) - sold;
      e.at(k, N_UNF) = tot - sold;
    }
  }
  for (int k = 0; k < sp.nf; ++k) {
    T unf = T(0);
    for (int i = sp.cust_start[k]; i < sp.cust_start[k + 1]; ++i) unf = unf + e.at(sp.cust_list[i], N_B);
    unf = unf + e.at(k, N_UNF);
    const double hk = -(double)e.at(k, N_INV) * sp.h[k], bk = -(double)unf * sp.b[k];
    ch += hk;
    cb += bk;
    if (cost != nullptr) {
      cost[(int64_t)k * ld + env] = hk;
      cost[(int64_t)(kNF + k) * ld + env] = bk;
    }
  }
  net_set_demand(e, sp, demand_next, ld, env);                                       // Node.update_demand
  if (!terminal) {
    // before_action (supplynetwork.py:185-199)
    for (int k = 0; k < sp.nf; ++k) {
      const int li = sp.li[k];
      const T q = e.at(k, N_Q);
      const T arrived = li == 1 ? q : e.at(k, N_INFO);
      e.at(k, N_INFO) = e.at(k, N_INFO + 1); e.at(k, N_INFO + 1) = e.at(k, N_INFO + 2); e.at(k, N_INFO + 2) = T(0);
      if (li >= 2) e.at(k, N_INFO + li - 2) = q;
      e.at(k, N_B) = e.at(k, N_B) + arrived;
      e.at(k, N_U + 3) = e.at(k, N_U + 3) + e.at(k, N_U4);
      e.at(k, N_Q) = arrived;                           // kept for the latest-demand pass below
    }
    for (int k = 0; k < sp.nf; ++k) e.at(k, N_LATI) = sp.lat_src[k] >= 0 ? e.at(sp.lat_src[k], N_Q) : T(0);
  }
  return ch + cb;
}

// observation (supplynetwork.py:151-169 per node, envs.py:103-110): node k lands at dst + 7*pos[k]
template <typename T>
__device__ __forceinline__ void net_emit_obs(const NetEnv<T>& e, const NetSpec<T>& sp, bool terminal, float* __restrict__ dst) {
  for (int k = 0; k < sp.nf; ++k) {
    if (sp.pos[k] < 0) continue;
    T u0 = e.at(k, N_U), u1 = e.at(k, N_U + 1), u2 = e.at(k, N_U + 2), u3 = e.at(k, N_U + 3);
    const T unf = e.unfilled(sp, k), lat = e.latest(k);
    if (sp.pre[k] && !terminal) {          // already ordered inside before_action: 5-slot list, first four read
      const T q = tmax(e.base_stock_order(sp, k, sp.target[k], sp.clb, sp.cub, 0), T(0));
      u3 = u2; u2 = u1; u1 = u0; u0 = q;
    }
    float* d = dst + 7 * sp.pos[k];
    d[0] = (float)e.at(k, N_INV); d[1] = (float)unf; d[2] = (float)lat;
    d[3] = (float)u0; d[4] = (float)u1; d[5] = (float)u2; d[6] = (float)u3;
  }
}

// Row-major observations through a per-warp staging buffer in shared memory: every lane builds its row there, then the
// warp copies the 32 consecutive rows to global memory with fully coalesced stores (a lane writing its own 7*n_obs
// floats directly would touch 32 different lines per store instruction).  All 32 lanes must call this.
//   stage      this warp's buffer, 32 * 7 * n_obs floats
//   dst        global address of the row of the warp's first env
//   n_valid    number of real envs in this warp (lanes >= n_valid only help copying)
template <typename T>
__device__ __forceinline__ void net_write_obs(const NetEnv<T>& e, const NetSpec<T>& sp, bool terminal, float* stage,
                                              float* __restrict__ dst, int n_valid) {
  const int lane = threadIdx.x & 31, od = 7 * sp.n_obs;
  if (lane < n_valid) net_emit_obs(e, sp, terminal, stage + lane * od);
  __syncwarp();
  const int total = n_valid * od;
  for (int i = lane; i < total; i += 32) dst[i] = stage[i];
  __syncwarp();
}

}  // namespace sn
