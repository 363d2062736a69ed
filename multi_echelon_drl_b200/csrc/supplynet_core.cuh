// supplynet_core.cuh — per-env beer-game recurrence held entirely in registers.
//
// One `Env<T>` is the collapsed state of one reference SupplyNetwork (supplynetwork.py:10-343
// over network_components.py:8-383) for the serial 4-echelon chain: the Order/Shipment object
// lists become fixed-size pipelines because lead times are fixed and `Arc.fill_orders`
// (network_components.py:203-234) ships min(inventory, backlog) in aggregate.
//
// State word map (row index into the [SN_STATE_WORDS][ld] structure-of-arrays buffer):
//   7k+0  inv[k]          on-hand inventory of node k            (supplynetwork.py:123)
//   7k+1  k=0: unf_ind    unfilled independent demand of the retailer; the observed
//                         "unfilled_demand" is lat[0] + unf_ind                     (:130-136)
//         k>0: B[k-1]     backlog of arc k-1 = observed "unfilled_demand" of node k
//   7k+2  lat[k]          "latest_demand": k=0: current demand; k>0: order that reached node k
//                         at the last before_action                                 (:139)
//   7k+3..6 U[k][0..3]    unreceived pipeline of arc k, newest first               (:165-167)
//   28+k  info[k], k<3    order slip of arc k still one period away from its supplier
//   31+2k+j ship[k][j]    shipment on arc k arriving after j+1 more sweeps
//   39    B[3]            backlog at the external supplier (arc 3)
// The first 28 words are the "global" observation of the reference (envs.py:103-110) except
// word 1, which the observation reports as lat[0] + unf_ind (kept apart so that fractional
// quantities accumulate exactly as Node.fill_independent_demand does).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <type_traits>

namespace sn {

// Facilities per env held by one thread.  4 everywhere except in a library specialised for a tree of five to eight
// nodes (specialize.py builds it with -DSN_KF=8: the tree code is written against kF, the chain code is not
// instantiated there).
#ifndef SN_KF
#define SN_KF 4
#endif
constexpr int kF = SN_KF;
constexpr int kObsPerNode = 7;
constexpr int kStateWords = 40;
constexpr int kInitWords = 19;
constexpr int kStatsWords = 16;
constexpr int kMaxItems = 4;

template <typename T>
struct DevSpec {
  T target[kF];        // co-player base-stock targets
  T clb, cub;          // co-player clip (utils/heuristics.py:16)
  T off, lb, ub;       // d+a rule (utils/wrappers.py:25-27)
  double h[kMaxItems][kF], b[kMaxItems][kF];   // per item (chain c belongs to item c % n_items)
  int n_items;
  int li[kF], ls[kF];  // information / shipment lead time of arc k (generic path, supplynet_generic.cuh)
  int nf;              // real facilities (1..4).  Chains shorter than 4 run on the generic path with the missing
                       // upstream nodes as phantoms: stock `big`, zero costs, never ordering, never observed, so
                       // that node nf behaves exactly like the unlimited external supplier of arc nf-1
  T big;
  int slot[kF];        // action column of node k, -1 = co-player
  int pos[kF];         // position of node k's 7 numbers in the observation, -1 = not observed
  int n_obs;
  int n_agents;
  int lo;              // first agent index: nodes below it order inside before_action
  int rule;
  int t_max;
  // distribution trees (supplynet_tree.cuh); chains leave these at their chain values
  int sup[kF];         // supplier node of node k, -1 = external supplier (also for phantom nodes)
  int rank[kF];        // rank of node k among its supplier's customers (served in this order)
  int max_rank;        // largest rank in use (rounds of the fill loop)
  int lat_src[kF];     // customer whose arrived order is node j's latest demand (the customer latest in the order
                       // sequence overwrites the others, supplynetwork.py:193), -1 = none
  int src[kF];         // row of node k's demand stream in the demand tables, -1 = not a demand source
  int n_src;
  int pre[kF];         // node k orders inside before_action (its place in the order sequence is ahead of the first agent)
  // the same tables as exact 0/1 multipliers of the quantity type: the generic / tree step applies a run-time
  // position as ONE multiply-add with a constant-bank operand (instead of compare + select + multiply)
  T m_ls[kF][4];       // 1 at slot ls[k]-1: where this period's shipment enters the pipeline
  T m_li[kF][3];       // 1 at slot li[k]-2: where this period's order slip enters the pipeline
  T m_li1[kF];         // 1 if li[k] == 1: the slip placed now is the one that arrives now
  T m_sup[kF][kF];     // [k][j] = 1 if node j supplies node k
  T m_ext[kF];         // 1 if node k buys from the external supplier
  T m_rank[kF - 1][kF];// [r][k] = 1 if node k has rank r among its supplier's customers
  T m_lat[kF][kF];     // [j][k] = 1 if lat_src[j] == k
};

template <typename T>
struct DevPolicy {
  T target[kF];        // indexed by NODE (the host permutes the per-column targets)
  T lb, ub;
  int rule;
};

template <typename T>
struct Env {
  T inv[kF], lat[kF], U[kF][4], B[kF], info[3], ship[kF][2], unf_ind;

  // "unfilled_demand" of node k as SupplyNetwork.get_state reports it (supplynetwork.py:130-136)
  __device__ __forceinline__ T unfilled(int k, const DevSpec<T>&) const { return k == 0 ? lat[0] + unf_ind : B[k - 1]; }
  __device__ __forceinline__ T latest(int k) const { return lat[k]; }                       // supplynetwork.py:139
  // co-player that orders inside before_action (chain position below the first agent)
  static __device__ __forceinline__ bool pre_agent(int k, const DevSpec<T>& sp) { return k < kF - 1 && k < sp.lo; }
  // words outside inv[] / B[] that enter sums unclipped (range guard of the integer path, see qty_mask)
  __device__ __forceinline__ uint32_t extra_qty_mask() const { return (uint32_t)unf_ind | (uint32_t)lat[0]; }
};

template <typename T> __device__ __forceinline__ T tmin(T a, T b) { return a < b ? a : b; }
template <typename T> __device__ __forceinline__ T tmax(T a, T b) { return a > b ? a : b; }
template <> __device__ __forceinline__ float tmin<float>(float a, float b) { return fminf(a, b); }
template <> __device__ __forceinline__ float tmax<float>(float a, float b) { return fmaxf(a, b); }
template <> __device__ __forceinline__ double tmin<double>(double a, double b) { return fmin(a, b); }
template <> __device__ __forceinline__ double tmax<double>(double a, double b) { return fmax(a, b); }
template <typename T> __device__ __forceinline__ T tclamp(T x, T lo, T hi) { return tmin(tmax(x, lo), hi); }

// utils/heuristics.py:43-57,60-75: target - (on_hand + on_order - unfilled) [- latest], clipped.
// on_order is summed slot 0..3 in that order, as both reference paths do.
template <typename T>
__device__ __forceinline__ T base_stock(T target, T inv, const T (&U)[4], T unf, T lat, T lb, T ub, int rule) {
  T on_order = ((U[0] + U[1]) + U[2]) + U[3];
  T q = target - (inv + on_order - unf);
  if (rule == 1) q = q - lat;
  return tclamp(q, lb, ub);
}

// Pack / unpack between the word map above and registers ------------------------------------
template <typename T>
__device__ __forceinline__ void env_from_words(Env<T>& e, const T (&w)[kStateWords]) {
#pragma unroll
  for (int k = 0; k < kF; ++k) {
    e.inv[k] = w[7 * k];
    if (k == 0) e.unf_ind = w[1]; else e.B[k - 1] = w[7 * k + 1];
    e.lat[k] = w[7 * k + 2];
#pragma unroll
    for (int j = 0; j < 4; ++j) e.U[k][j] = w[7 * k + 3 + j];
    e.ship[k][0] = w[31 + 2 * k];
    e.ship[k][1] = w[32 + 2 * k];
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) e.info[k] = w[28 + k];
  e.B[3] = w[39];
}

template <typename T>
__device__ __forceinline__ void env_to_words(const Env<T>& e, T (&w)[kStateWords]) {
#pragma unroll
  for (int k = 0; k < kF; ++k) {
    w[7 * k] = e.inv[k];
    w[7 * k + 1] = (k == 0) ? e.unf_ind : e.B[k - 1];
    w[7 * k + 2] = e.lat[k];
#pragma unroll
    for (int j = 0; j < 4; ++j) w[7 * k + 3 + j] = e.U[k][j];
    w[31 + 2 * k] = e.ship[k][0];
    w[32 + 2 * k] = e.ship[k][1];
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) w[28 + k] = e.info[k];
  w[39] = e.B[3];
}

// reset: Arc.reset / Node.reset (network_components.py:138-168, 293-319) + before_action(0).
// init words: inv[0..3]; so[k][j] at 4 + {0,2,4,6}[k] + j; sh[k][j] at 11 + 2k + j.
template <typename T>
__device__ __forceinline__ void env_reset(Env<T>& e, const T (&in)[kInitWords], T d0) {
  const int so_off[kF] = {4, 6, 8, 10};
#pragma unroll
  for (int k = 0; k < kF; ++k) {
    const T so0 = in[so_off[k]];
    const T so1 = (k < 3) ? in[so_off[k] + 1] : T(0);
    const T sh0 = in[11 + 2 * k], sh1 = in[12 + 2 * k];
    e.inv[k] = in[k];
    // U = ([0]*4 + shipments + sales_orders)[::-1][:5]   (network_components.py:168), then
    // before_action pops the last slot into its neighbour (supplynetwork.py:198-199).
    if (k < 3) {            // [so1, so0, sh1, sh0, 0] -> [so1, so0, sh1, sh0]
      e.U[k][0] = so1; e.U[k][1] = so0; e.U[k][2] = sh1; e.U[k][3] = sh0;
      e.info[k] = so1;      // so0 reaches the supplier in before_action(0)
    } else {                // [so0, sh1, sh0, 0, 0] -> [so0, sh1, sh0, 0]
      e.U[k][0] = so0; e.U[k][1] = sh1; e.U[k][2] = sh0; e.U[k][3] = T(0);
    }
    e.ship[k][0] = sh0; e.ship[k][1] = sh1;
    // order slip with remaining lead time 1 arrives now: backlog and latest demand of the supplier
    e.B[k] = so0;
    if (k < 3) e.lat[k + 1] = so0;
  }
  e.unf_ind = T(0);
  e.lat[0] = d0;            // current_external_demand (network_components.py:317-319)
}

// Observation of node K (supplynetwork.py:151-169), K a compile-time constant so that every
// operand is a fixed register.  A co-player with index < first agent has already ordered inside
// before_action when the reference observes it, so (on non-terminal observations) its pipeline
// shows [its order, U0, U1, U2] (5-slot list, first four read).
template <int K, typename T, typename E>
__device__ __forceinline__ void node_view(const E& e, const DevSpec<T>& sp, bool terminal, float (&v)[7]) {
  T u0 = e.U[K][0], u1 = e.U[K][1], u2 = e.U[K][2], u3 = e.U[K][3];
  const T unf = e.unfilled(K, sp), lat = e.latest(K);
  if (E::pre_agent(K, sp) && !terminal) {
    const T q = tmax(base_stock(sp.target[K], e.inv[K], e.U[K], unf, lat, sp.clb, sp.cub, 0), T(0));
    u3 = u2; u2 = u1; u1 = u0; u0 = q;
  }
  v[0] = (float)e.inv[K]; v[1] = (float)unf; v[2] = (float)lat;
  v[3] = (float)u0; v[4] = (float)u1; v[5] = (float)u2; v[6] = (float)u3;
}

// envs.py:103-110: concatenation over internal nodes (global) or agent-managed facilities, written
// to `dst` (this env's row, shared or global memory).  Node K lands at dst + 7*pos[K]: the position
// is a memory offset, never a register select.  IDENT = all four nodes in chain order (the
// centralized / full-information layout): 28 floats as seven 16-byte stores.
template <bool IDENT, typename T, typename E>
__device__ __forceinline__ void emit_obs(const E& e, const DevSpec<T>& sp, bool terminal, float* dst) {
  if (IDENT) {
    float o[28];
    { float v[7]; node_view<0>(e, sp, terminal, v);
#pragma unroll
      for (int j = 0; j < 7; ++j) o[j] = v[j]; }
    { float v[7]; node_view<1>(e, sp, terminal, v);
#pragma unroll
      for (int j = 0; j < 7; ++j) o[7 + j] = v[j]; }
    { float v[7]; node_view<2>(e, sp, terminal, v);
#pragma unroll
      for (int j = 0; j < 7; ++j) o[14 + j] = v[j]; }
    { float v[7]; node_view<3>(e, sp, terminal, v);
#pragma unroll
      for (int j = 0; j < 7; ++j) o[21 + j] = v[j]; }
    if ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0) {
#pragma unroll
      for (int j = 0; j < 7; ++j)
        reinterpret_cast<float4*>(dst)[j] = make_float4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
    } else {                       // caller's buffer is not 16-byte aligned (plain-store path only)
#pragma unroll
      for (int j = 0; j < 28; ++j) dst[j] = o[j];
    }
  } else {
#define SN_EMIT_NODE(K)                                        \
    if (sp.pos[K] >= 0) {                                      \
      float v[7];                                              \
      node_view<K>(e, sp, terminal, v);                        \
      float* d = dst + 7 * sp.pos[K];                          \
      _Pragma("unroll") for (int j = 0; j < 7; ++j) d[j] = v[j]; \
    }
    SN_EMIT_NODE(0) SN_EMIT_NODE(1) SN_EMIT_NODE(2) SN_EMIT_NODE(3)
#if SN_KF > 4
    SN_EMIT_NODE(4) SN_EMIT_NODE(5) SN_EMIT_NODE(6) SN_EMIT_NODE(7)
#endif
#undef SN_EMIT_NODE
  }
}

// Range guard of the integer path.  The reference computes with Python ints (unbounded); int32 rollouts equal it
// exactly while no intermediate wraps, which holds as long as every order and every stock / backlog stays below
// 2^26 (67,108,864): pipelines and on-order sums are sums of at most a handful of such terms.  The kernels report
// (never correct) an env-step whose action magnitude, order, inventory or backlog reached that bound: a sticky count
// in stats[13] / bit 1 of ctl[2].  Floating-point paths have no such limit.
constexpr int kQtyLimitLog2 = 26;
// OR of the magnitudes to be watched: the raw actions and the orders placed (taken BEFORE the step, so that the
// action registers can die there; not needed when both come from a clipped in-kernel policy) and, after the step, the
// stocks and backlogs.  The remaining words are bounded by these: pipelines hold earlier orders, shipments are
// min(inventory, backlog), on-order sums have at most eight such terms.
template <typename T>
__device__ __forceinline__ uint32_t qty_mask_orders(const T (&a)[kF], const T (&q)[kF]) {
  if constexpr (std::is_same<T, int32_t>::value) {
    uint32_t m = 0;
#pragma unroll
    for (int k = 0; k < kF; ++k) m |= (uint32_t)q[k] | (uint32_t)(a[k] ^ (a[k] >> 31));
    return m;
  } else {
    return 0u;
  }
}
template <typename T, typename E>
__device__ __forceinline__ uint32_t qty_mask_state(const E& e) {
  if constexpr (std::is_same<T, int32_t>::value) {
    uint32_t m = e.extra_qty_mask();
#pragma unroll
    for (int k = 0; k < kF; ++k) m |= (uint32_t)e.inv[k] | (uint32_t)e.B[k];
    return m;
  } else {
    return 0u;
  }
}
#ifdef SN_NO_RANGE_GUARD                     /* experiments only: what the guard costs */
__device__ __forceinline__ bool qty_mask_over(uint32_t) { return false; }
#else
__device__ __forceinline__ bool qty_mask_over(uint32_t m) { return (m >> kQtyLimitLog2) != 0; }
#endif
struct StepOut {
  double reward;
  double hold[kF], back[kF];
};

// item of the chain `env` (multi-item batches interleave the items, see sn_spec); 0 for single-item
template <typename T>
__device__ __forceinline__ int chain_item(const DevSpec<T>& sp, int64_t env) {
  return sp.n_items > 1 ? (int)((uint32_t)env % (uint32_t)sp.n_items) : 0;
}

// Orders of all four nodes for this period.  Each depends only on start-of-period quantities
// (every information lead time is >= 1), so they are computed before the shipment sweep:
// agent-managed nodes take their action (through the d+a rule of utils/wrappers.py:25-27 when
// selected), the others their own base-stock policy (network_components.py:361-366).
//   a[k]  action of NODE k (already gathered from its action column); ignored for co-players
template <bool IDENT, typename T, typename E>
__device__ __forceinline__ void env_orders(const E& e, const DevSpec<T>& sp, const T (&a)[kF], T (&q)[kF]) {
#pragma unroll
  for (int k = 0; k < kF; ++k) {
    T x;
    if (IDENT || sp.slot[k] >= 0) {
      x = a[k];
      if (sp.rule == 1) x = tclamp(e.latest(k) + x + sp.off, sp.lb, sp.ub);
    } else {
      x = base_stock(sp.target[k], e.inv[k], e.U[k], e.unfilled(k, sp), e.latest(k), sp.clb, sp.cub, 0);
    }
    q[k] = tmax(x, T(0));                                                  // network_components.py:368-374
  }
}

// One period: agent_action + after_action + get_cost + (before_action | terminal observation).
//   q[k]      order placed by node k this period (env_orders)
//   d_next    demand of the next period
//   terminal  period + 1 >= max_episode_steps
// On a terminal step `e` becomes the *terminal observation view* (envs.py:98-110: no
// before_action, 5-slot pipelines of which the first four are read, stale latest demand) and is
// not a valid state any more.
template <bool MULTI, typename T>
__device__ __forceinline__ void env_step(Env<T>& e, const DevSpec<T>& sp, int item, const T (&q)[kF], T d_next,
                                         bool terminal, StepOut& out) {
  // ---- after_action sweep, upstream first (supplynetwork.py:241-264)
  T U4[kF];   // fifth (oldest) slot of [q, U0, U1, U2, U3] after this period's arrivals
#pragma unroll
  for (int k = kF - 1; k >= 0; --k) {
    T u[5] = {q[k], e.U[k][0], e.U[k][1], e.U[k][2], e.U[k][3]};           // U.insert(0, q)
    const T arr = e.ship[k][0];                                            // advance_and_receive_shipments
    e.ship[k][0] = e.ship[k][1];
    e.inv[k] = e.inv[k] + arr;
    T rem = arr;
#pragma unroll
    for (int i = 4; i >= 0; --i) {                                         // :250-255, oldest slot first
      const T f = tmin(u[i], rem);
      u[i] = u[i] - f;
      rem = rem - f;
    }
    T s;
    if (k == kF - 1) {
      s = e.B[k];                                                          // external supplier: unlimited
    } else {
      s = tmin(e.inv[k + 1], e.B[k]);                                      // Arc.fill_orders in aggregate
      e.inv[k + 1] = e.inv[k + 1] - s;
    }
    e.ship[k][1] = s;
    e.B[k] = e.B[k] - s;
    e.U[k][0] = u[0]; e.U[k][1] = u[1]; e.U[k][2] = u[2]; e.U[k][3] = u[3];
    U4[k] = u[4];
  }
  // retailer fills independent demand (network_components.py:323-345)
  const T tot = e.lat[0] + e.unf_ind;
  const T sold = tmax(T(0), tmin(e.inv[0], tot));
  e.inv[0] = e.inv[0] - sold;
  e.unf_ind = tot - sold;

  // ---- get_cost (supplynetwork.py:285-310): node-list order, c_h and c_b accumulated apart.  Explicitly rounded
  // products and sums (__dmul_rn / __dadd_rn are never contracted into fused multiply-adds): the reference rounds
  // after every operation, and with cost coefficients that are not dyadic a fused multiply-add would differ
  double ch = 0.0, cb = 0.0;
  if (MULTI) {                     // per-chain cost vector (dynamic constant-bank index)
#pragma unroll
    for (int k = 0; k < kF; ++k) {
      const T unfk = (k == 0) ? e.unf_ind : e.B[k - 1];
      out.hold[k] = __dmul_rn(-(double)e.inv[k], sp.h[item][k]);
      out.back[k] = __dmul_rn(-(double)unfk, sp.b[item][k]);
      ch = __dadd_rn(ch, out.hold[k]);
      cb = __dadd_rn(cb, out.back[k]);
    }
  } else {                         // single item: coefficients fold into the multiply as constant operands
#pragma unroll
    for (int k = 0; k < kF; ++k) {
      const T unfk = (k == 0) ? e.unf_ind : e.B[k - 1];
      out.hold[k] = __dmul_rn(-(double)e.inv[k], sp.h[0][k]);
      out.back[k] = __dmul_rn(-(double)unfk, sp.b[0][k]);
      ch = __dadd_rn(ch, out.hold[k]);
      cb = __dadd_rn(cb, out.back[k]);
    }
  }
  out.reward = __dadd_rn(ch, cb);

  e.lat[0] = d_next;                                                       // Node.update_demand (:266-268)
  // ---- next period (envs.py:96-101)
  if (!terminal) {
    // before_action (supplynetwork.py:185-199): slips advance, the arrived slip joins the backlog
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const T arrived = e.info[k];
      e.info[k] = q[k];
      e.lat[k + 1] = arrived;
      e.B[k] = e.B[k] + arrived;
    }
    e.B[3] = e.B[3] + q[3];                   // info lead time 1: this period's order arrives now
#pragma unroll
    for (int k = 0; k < kF; ++k) e.U[k][3] = e.U[k][3] + U4[k];            // pop last, merge (:198-199)
  }
}

}  // namespace sn
