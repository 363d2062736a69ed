// supplynet_net.cuh — distribution trees of FIVE TO EIGHT internal nodes (continued).
//
// Same semantics as supplynet_tree.cuh (every node buys from one supplier, may serve several customers and / or its
// own demand; what SupplyNetwork does for such node / arc lists, supplynetwork.py:182-268), different decomposition:
// 8 nodes x 16 state words do not fit one thread's registers at a useful occupancy, so an env is spread over EIGHT
// LANES of a warp, one lane per node (four envs per warp).  A lane keeps its node's 16 words in registers and runs
// the same per-node code as the chain kernels; what crosses nodes travels by warp shuffles inside the 8-lane group:
//   * a customer reads its supplier's inventory (one shuffle) and the backlog of the customers served before it
//     (8 shuffles + a bit mask), and takes s = min(B, max(0, A - P)): the closed form of "the supplier serves its
//     customer arcs in rank order from what it has" (s_0 = min(A, B_0), s_1 = min(A - s_0, B_1), ...);
//   * a supplier subtracts what its customers took and sums their backlog for its observation / cost (8 shuffles);
//   * the latest demand of a node is the order that just reached it from one particular customer (one shuffle);
//   * the reward is summed over the nodes in node order (the reference's order) with 2 x 8 double shuffles.
// Observation rows are assembled per warp in shared memory (envs x 7 x n_obs floats) and leave coalesced.
// An env takes L lanes, L = 5, 6 or 8 by the node count (template constant): six envs of five nodes, five of six or
// four of seven / eight share a warp (the two spare lanes of the first two cases idle), and every loop over the
// nodes of an env is L long.  (Seven lanes for seven nodes: the same four envs per warp and no butterfly for the
// reward sum: measured slower than eight, 0.758 vs 0.719 ms.)
// An earlier version kept each env in shared memory under one thread (profiles/r1_notes.md): 4-5x slower.
//
// State word map in HBM ([SN_NET_STATE_WORDS = 128][ld]; node k owns words 16k .. 16k+15):
//   +0 inv   +1 unf_ind   +2 latI (latest internal demand)   +3..6 U[0..3]   +7..9 info[0..2]   +10..13 ship[0..3]
//   +14 B (backlog of the node's upstream arc)   +15 dem (current external demand)
// Init words ([SN_NET_INIT_WORDS = 72][ld]): inv[8]; so[k][j] at 8+4k+j; sh[k][j] at 40+4k+j (unused = 0).
// Demand tables: one row per demand source and period, as for the smaller trees.
#pragma once
#include "supplynet_core.cuh"

namespace sn {

constexpr int kNF = 8;                        // most nodes per env (array extents; lanes per env: L <= kNF)
constexpr int kNetMaxEnvsPerWarp = 6;         // L = 5
constexpr int kNetStateWords = 16 * kNF;      // 128
constexpr int kNetInitWords = kNF + 8 * kNF;  // 72
constexpr int kNetBlock = 128;                // 4 warps = 16 .. 24 envs per block
__host__ __device__ constexpr int net_lanes(int nf) { return nf <= 5 ? 5 : nf == 6 ? 6 : 8; }
constexpr unsigned kFullMask = 0xffffffffu;

template <typename T>
struct NetSpec {
  int nf, n_agents, n_obs, n_src, rule, t_max;
  int exact_costs;                 // cost sums cannot round (see group_reward)
  T clb, cub, off, lb, ub, big;
  T target[kNF];
  double h[kNF], b[kNF];
  int li[kNF], ls[kNF];
  int sup[kNF];                    // supplier node, -1 = external supplier
  int slot[kNF], pos[kNF];         // action column / observation position of node k, -1 = none
  int src[kNF];                    // demand-table row of node k, -1 = no external demand
  int pre[kNF];                    // orders inside before_action (ahead of the first agent in the order sequence)
  int lat_src[kNF];                // customer whose arriving order is node k's latest demand, -1 = none
  unsigned cust[kNF];              // bit j: node j is a customer of node k
  unsigned before[kNF];            // bit j: node j has the same (internal) supplier as k and is served before it
};

template <typename T>
struct NetPolicy {
  T target[kNF];                   // per NODE
  T lb, ub;
  int rule;
};

// what a lane needs to know about its node
template <typename T>
struct LaneCfg {
  bool active;                     // a real node of a real env
  int k, base;                     // node index; first lane of this env's group
  int li, ls, sup, slot, pos, src, pre, lat_src;
  unsigned cust, before;
  T cust_m[kNF], before_m[kNF];    // the same two bit masks as exact 0/1 multipliers: one multiply-add per node and sum
  T ship_m[4], info_m[3], lat_m[kNF], li1_m;   // 1 at slot ls-1 / slot li-2 / node lat_src / if li == 1
  T target, pol_target;
  double h, b;
};

template <int L, typename T>
__device__ __forceinline__ LaneCfg<T> lane_cfg(const NetSpec<T>& sp, const NetPolicy<T>& pol, bool valid_env) {
  LaneCfg<T> c;
  const int lane = threadIdx.x & 31;
  c.k = lane % L;
  c.base = lane - c.k;
  c.active = valid_env && c.k < sp.nf;               // (valid_env is false on the spare lanes of a warp)
  c.li = sp.li[c.k]; c.ls = sp.ls[c.k]; c.sup = sp.sup[c.k]; c.slot = sp.slot[c.k]; c.pos = sp.pos[c.k];
  c.src = sp.src[c.k]; c.pre = sp.pre[c.k]; c.lat_src = sp.lat_src[c.k];
  c.cust = sp.cust[c.k]; c.before = sp.before[c.k];
#pragma unroll
  for (int j = 0; j < kNF; ++j) {
    c.cust_m[j] = (T)((c.cust >> j) & 1u);
    c.before_m[j] = (T)((c.before >> j) & 1u);
    c.lat_m[j] = (T)(j == c.lat_src);
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) c.ship_m[j] = (T)(j == c.ls - 1);
#pragma unroll
  for (int j = 0; j < 3; ++j) c.info_m[j] = (T)(j == c.li - 2);
  c.li1_m = (T)(c.li == 1);
  c.target = sp.target[c.k]; c.pol_target = pol.target[c.k];
  c.h = sp.h[c.k]; c.b = sp.b[c.k];
  return c;
}

// one node of one env; inactive lanes hold zeros
template <typename T>
struct LaneEnv {
  T inv, unf, latI, U[4], info[3], ship[4], B, dem;
};

template <typename T>
__device__ __forceinline__ void lane_zero(LaneEnv<T>& e) {
  e.inv = e.unf = e.latI = e.B = e.dem = T(0);
#pragma unroll
  for (int j = 0; j < 4; ++j) e.U[j] = e.ship[j] = T(0);
#pragma unroll
  for (int j = 0; j < 3; ++j) e.info[j] = T(0);
}

// sum of `v` over the lanes of this env whose bit is set in `mask` (all 32 lanes take part)
template <int L, typename T>
__device__ __forceinline__ T group_sum(T v, unsigned mask, int base) {
  T s = T(0);
#pragma unroll
  for (int j = 0; j < L; ++j) {
    const T x = __shfl_sync(kFullMask, v, base + j);
    s = s + x * (T)((mask >> j) & 1u);
  }
  return s;
}

// the values of `v` held by the nodes of this env (entries L.. of `out` are never touched nor read)
template <int L, typename T>
__device__ __forceinline__ void group_all(T v, int base, T (&out)[kNF]) {
#pragma unroll
  for (int j = 0; j < L; ++j) out[j] = __shfl_sync(kFullMask, v, base + j);
}

// sum over the nodes whose bit is set in `mask` of values every lane already holds
template <int L, typename T>
__device__ __forceinline__ T masked_sum(const T (&v)[kNF], unsigned mask) {
  T s = T(0);
#pragma unroll
  for (int j = 0; j < L; ++j) s = s + v[j] * (T)((mask >> j) & 1u);
  return s;
}

template <int L, typename T>
__device__ __forceinline__ T weighted_sum(const T (&v)[kNF], const T (&w)[kNF]) {
  T s = T(0);
#pragma unroll
  for (int j = 0; j < L; ++j) s = s + v[j] * w[j];
  return s;
}

// value of `v` held by node `node` of this env (node < 0: lane 0 of the group is read and must be ignored)
template <typename T>
__device__ __forceinline__ T group_get(T v, int node, int base) {
  return __shfl_sync(kFullMask, v, base + (node < 0 ? 0 : node));
}

template <typename T>
__device__ __forceinline__ void lane_load(LaneEnv<T>& e, const LaneCfg<T>& c, const T* __restrict__ state, int64_t ld, int64_t env) {
  lane_zero(e);
  if (!c.active) return;
  const T* p = state + (int64_t)(16 * c.k) * ld + env;
  e.inv = p[0]; e.unf = p[ld]; e.latI = p[2 * ld];
#pragma unroll
  for (int j = 0; j < 4; ++j) e.U[j] = p[(3 + j) * ld];
#pragma unroll
  for (int j = 0; j < 3; ++j) e.info[j] = p[(7 + j) * ld];
#pragma unroll
  for (int j = 0; j < 4; ++j) e.ship[j] = p[(10 + j) * ld];
  e.B = p[14 * ld]; e.dem = p[15 * ld];
}

template <typename T>
__device__ __forceinline__ void lane_store(const LaneEnv<T>& e, const LaneCfg<T>& c, T* __restrict__ state, int64_t ld, int64_t env) {
  if (!c.active) return;
  T* p = state + (int64_t)(16 * c.k) * ld + env;
  p[0] = e.inv; p[ld] = e.unf; p[2 * ld] = e.latI;
#pragma unroll
  for (int j = 0; j < 4; ++j) p[(3 + j) * ld] = e.U[j];
#pragma unroll
  for (int j = 0; j < 3; ++j) p[(7 + j) * ld] = e.info[j];
#pragma unroll
  for (int j = 0; j < 4; ++j) p[(10 + j) * ld] = e.ship[j];
  p[14 * ld] = e.B; p[15 * ld] = e.dem;
}

// external demand of this lane's node in one period's rows of the demand table (0 without a stream)
template <typename T>
__device__ __forceinline__ T lane_demand(const LaneCfg<T>& c, const T* __restrict__ row, int64_t ld, int64_t env) {
  return (c.active && c.src >= 0) ? __ldg(row + (int64_t)c.src * ld + env) : T(0);
}

// Arc.reset / Node.reset + before_action(0) (network_components.py:138-168,293-319; supplynetwork.py:182-199)
template <typename T>
__device__ __forceinline__ void lane_reset(LaneEnv<T>& e, const LaneCfg<T>& c, const T* __restrict__ init, int64_t ld,
                                           int64_t env, const T* __restrict__ demand0) {
  lane_zero(e);
  T so[4] = {T(0), T(0), T(0), T(0)}, sh[4] = {T(0), T(0), T(0), T(0)};
  if (c.active) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (j < c.li) so[j] = __ldg(init + (int64_t)(kNF + 4 * c.k + j) * ld + env);
      if (j < c.ls) sh[j] = __ldg(init + (int64_t)(5 * kNF + 4 * c.k + j) * ld + env);
    }
    e.inv = __ldg(init + (int64_t)c.k * ld + env);
  }
  // ([0]*4 + shipments + sales_orders)[::-1][:5]: newest slip first, then shipments latest first
  T seq[5];
#pragma unroll
  for (int i = 0; i < 5; ++i) {
    T v = T(0);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      v = v + so[j] * (T)(c.li - 1 - i == j);
      v = v + sh[j] * (T)(i >= c.li && c.ls - 1 - (i - c.li) == j);
    }
    seq[i] = v;
  }
  e.U[0] = seq[0]; e.U[1] = seq[1]; e.U[2] = seq[2]; e.U[3] = seq[3] + seq[4];
  e.B = so[0];                                             // the slip with remaining lead time 1 arrives now
  e.info[0] = so[1]; e.info[1] = so[2]; e.info[2] = so[3];
#pragma unroll
  for (int j = 0; j < 4; ++j) e.ship[j] = sh[j];
  const T from = group_get(so[0], c.lat_src, c.base);
  e.latI = c.lat_src >= 0 ? from : T(0);
  e.dem = lane_demand(c, demand0, ld, env);
}

// observed unfilled_demand / latest_demand (supplynetwork.py:130-139); `cust_backlog` = group_sum(B, cust)
template <typename T>
__device__ __forceinline__ T lane_unfilled(const LaneEnv<T>& e, T cust_backlog) { return cust_backlog + (e.dem + e.unf); }
template <typename T>
__device__ __forceinline__ T lane_latest(const LaneEnv<T>& e) { return e.latI + e.dem; }

// this node's order for the period (depends on start-of-period quantities only)
template <int POLICY, typename T>
__device__ __forceinline__ T lane_order(const LaneEnv<T>& e, const LaneCfg<T>& c, const NetSpec<T>& sp, const NetPolicy<T>& pol,
                                        T cust_backlog, T action, T& action_taken) {
  const T unf = lane_unfilled(e, cust_backlog), lat = lane_latest(e);
  T x;
  if (c.slot >= 0) {
    x = (POLICY == SN_POLICY_BASE_STOCK) ? base_stock(c.pol_target, e.inv, e.U, unf, lat, pol.lb, pol.ub, pol.rule) : action;
    action_taken = x;
    if (sp.rule == 1) x = tclamp(lat + x + sp.off, sp.lb, sp.ub);                    // utils/wrappers.py:25-27
  } else {
    x = base_stock(c.target, e.inv, e.U, unf, lat, sp.clb, sp.cub, 0);                // network_components.py:361-366
  }
  return c.active ? tmax(x, T(0)) : T(0);                                             // :368-374
}

// One period of this lane's node.  `bj` holds the backlog of all eight upstream arcs of this env (every lane keeps the
// same copy: it is what customers-served-before-me and my-customers sums are taken from without further shuffles); it
// is updated to the new state.  Returns the period's holding / backorder cost terms of the node.
template <int L, typename T>
__device__ __forceinline__ void lane_step(LaneEnv<T>& e, const LaneCfg<T>& c, const NetSpec<T>& sp, T q, T d_next,
                                          bool terminal, T (&bj)[kNF], double& hold, double& back) {
  // receipt on the node's upstream arc (supplynetwork.py:243-255)
  T u[5] = {q, e.U[0], e.U[1], e.U[2], e.U[3]};
  const T arr = e.ship[0];
  e.inv = e.inv + arr;
  T rem = arr;
#pragma unroll
  for (int i = 4; i >= 0; --i) {
    const T f = tmin(u[i], rem);
    u[i] = u[i] - f;
    rem = rem - f;
  }
  e.ship[0] = e.ship[1]; e.ship[1] = e.ship[2]; e.ship[2] = e.ship[3]; e.ship[3] = T(0);
  // fills (Arc.fill_orders in aggregate, :260-262): closed form of the rank-ordered service, see the header
  const T a_sup = group_get(e.inv, c.sup, c.base);
  const T avail = c.sup < 0 ? sp.big : a_sup;
  const T amt = c.active ? tmin(e.B, tmax(T(0), avail - weighted_sum<L>(bj, c.before_m))) : T(0);
  T amt_j[kNF];
  group_all<L>(amt, c.base, amt_j);
#pragma unroll
  for (int j = 0; j < L; ++j) bj[j] = bj[j] - amt_j[j];
  e.inv = e.inv - weighted_sum<L>(amt_j, c.cust_m);
  e.B = e.B - amt;
#pragma unroll
  for (int j = 0; j < 4; ++j) e.ship[j] = e.ship[j] + amt * c.ship_m[j];
  // independent demand (network_components.py:323-345)
  if (c.src >= 0) {
    const T tot = e.dem + e.unf;
    const T sold = tmax(T(0), tmin(e.inv, tot));
    e.inv = e.inv - sold;
    e.unf = tot - sold;
  }
  // cost terms of this node (supplynetwork.py:285-310)
  const T unf_post = weighted_sum<L>(bj, c.cust_m) + e.unf;
  hold = -(double)e.inv * c.h;
  back = -(double)unf_post * c.b;
  e.dem = d_next;                                                                      // Node.update_demand
  e.U[0] = u[0]; e.U[1] = u[1]; e.U[2] = u[2];
  if (!terminal) {
    // before_action (supplynetwork.py:185-199)
    const T arrived = e.info[0] + q * c.li1_m;        // lead time 1: the head slot is empty, q arrives now
    e.info[0] = e.info[1]; e.info[1] = e.info[2]; e.info[2] = T(0);
#pragma unroll
    for (int j = 0; j < 3; ++j) e.info[j] = e.info[j] + q * c.info_m[j];
    e.B = e.B + arrived;
    e.U[3] = u[3] + u[4];
    T arr_j[kNF];
    group_all<L>(arrived, c.base, arr_j);
    T lat = T(0);
#pragma unroll
    for (int j = 0; j < L; ++j) {
      bj[j] = bj[j] + arr_j[j];
      lat = lat + arr_j[j] * c.lat_m[j];
    }
    e.latI = lat;
  } else {
    e.U[3] = u[3];                             // terminal view: 5-slot list of which the first four are read
  }
}

// reward = (sum of holding terms in node order) + (sum of backorder terms in node order), like get_cost.  When every
// term is an exact multiple of 2^-10 below 2^52 (integer quantities, cost coefficients that are multiples of 2^-10:
// NetSpec::exact_costs) no addition rounds, the order is immaterial and a 3-step butterfly does (12 shuffles
// instead of 32); otherwise the terms are added strictly in node order.
template <int L>
__device__ __forceinline__ double group_reward(double hold, double back, int base, bool exact) {
  if (L == 8 && exact) {                     // (the butterfly needs groups aligned to a power of two)
#pragma unroll
    for (int o = 1; o < kNF; o <<= 1) {
      hold += __shfl_xor_sync(kFullMask, hold, o);
      back += __shfl_xor_sync(kFullMask, back, o);
    }
    return hold + back;
  }
  double ch = 0.0, cb = 0.0;
#pragma unroll
  for (int j = 0; j < L; ++j) {
    ch += __shfl_sync(kFullMask, hold, base + j);
    cb += __shfl_sync(kFullMask, back, base + j);
  }
  return ch + cb;
}

// Observation rows of the warp's envs, assembled in `stage` ([32 / L][7*n_obs] floats) and copied out coalesced.
//   dst       global address of the row of the warp's first env;  n_valid: real envs in this warp (0 .. 32 / L)
template <int L, typename T>
__device__ __forceinline__ void lane_write_obs(const LaneEnv<T>& e, const LaneCfg<T>& c, const NetSpec<T>& sp, T cust_backlog,
                                               bool terminal, float* stage, float* __restrict__ dst, int n_valid) {
  const int lane = threadIdx.x & 31, od = 7 * sp.n_obs;
  if (c.active && c.pos >= 0) {
    const T unf = lane_unfilled(e, cust_backlog), lat = lane_latest(e);
    T u0 = e.U[0], u1 = e.U[1], u2 = e.U[2], u3 = e.U[3];
    if (c.pre && !terminal) {          // already ordered inside before_action: 5-slot list, first four read
      const T q = tmax(base_stock(c.target, e.inv, e.U, unf, lat, sp.clb, sp.cub, 0), T(0));
      u3 = u2; u2 = u1; u1 = u0; u0 = q;
    }
    float* d = stage + (lane / L) * od + 7 * c.pos;
    d[0] = (float)e.inv; d[1] = (float)unf; d[2] = (float)lat;
    d[3] = (float)u0; d[4] = (float)u1; d[5] = (float)u2; d[6] = (float)u3;
  }
  __syncwarp();
  const int total = n_valid * od;
  for (int i = lane; i < total; i += 32) dst[i] = stage[i];
  __syncwarp();
}

}  // namespace sn
