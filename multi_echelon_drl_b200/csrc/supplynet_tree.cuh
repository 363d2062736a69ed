// supplynet_tree.cuh — the same recurrence on DISTRIBUTION TREES: every internal node buys from
// one supplier (another node or the external supplier) and may serve several customers and / or its own external
// demand.  This is what the reference's SupplyNetwork does for such node / arc lists (supplynetwork.py:182-268):
//   * before_action walks the order sequence (utils/graph.py:5-32); the order slip that reaches a supplier
//     OVERWRITES its latest_demand (supplynetwork.py:193), so with several customers the one latest in the
//     order sequence is what the supplier "sees" (DevSpec::lat_src);
//   * after_action walks the reverse sequence: a node's customer arcs first receive, then the node serves them
//     in customers_dict order from what it has (the first arc has priority, DevSpec::rank), then its own demand;
//   * get_state: unfilled_demand = backlog of all customer arcs + current external demand + unfilled
//     independent demand (supplynetwork.py:130-136); get_cost: backlog of all customer arcs + unfilled independent.
// Receipts touch only the customer's inventory and pipeline, fills only the supplier's inventory and the arc's
// backlog, so the sweep is evaluated as: all receipts; fills in rounds by customer rank (different suppliers
// are independent); independent demand.  Supplier indices are applied with 0/1 mask arithmetic over the four
// nodes (registers only), like the lead times of supplynet_generic.cuh.
//
// State word map ([SN_TREE_STATE_WORDS = 64][ld]):
//   7k+0 inv[k]   7k+1 unf_ind[k]   7k+2 latI[k] (latest internal demand)   7k+3..6 U[k][0..3]
//   28+3k+j info[k][j]   40+4k+j ship[k][j]   56+k B[k] (backlog of node k's upstream arc)   60+k dem[k]
// Init words: as the generic path (36).  Demand tables: one row per demand source and period.
#pragma once
#include "supplynet_generic.cuh"

namespace sn {

constexpr int kTStateWords = 64;

template <typename T>
struct TreeDem { T v[kF]; };      // current external demand per NODE (0 for nodes without a stream)

template <typename T>
struct TEnv {
  T inv[kF], U[kF][4], B[kF], info[kF][kGI], ship[kF][kGS], unf_ind[kF], dem[kF], latI[kF];
  // observed unfilled_demand (supplynetwork.py:130-136): backlog of all customer arcs + (current external demand +
  // unfilled independent demand); observed latest_demand (:139).  Recomputed where read: registers are scarcer
  // than multiply-adds here.
  __device__ __forceinline__ T unfilled(int k, const DevSpec<T>& sp) const {
    T cb = T(0);
#pragma unroll
    for (int c = 0; c < kF; ++c) cb = cb + B[c] * sp.m_sup[c][k];
    return cb + (dem[k] + unf_ind[k]);
  }
  __device__ __forceinline__ T latest(int k) const { return latI[k] + dem[k]; }
  static __device__ __forceinline__ bool pre_agent(int k, const DevSpec<T>& sp) { return sp.pre[k] != 0; }
  __device__ __forceinline__ uint32_t extra_qty_mask() const {
    uint32_t m = 0;
#pragma unroll
    for (int k = 0; k < kF; ++k) m |= (uint32_t)unf_ind[k] | (uint32_t)dem[k];
    return m;
  }
};

// backlog of all customer arcs of node k
template <typename T>
__device__ __forceinline__ T customers_backlog(const TEnv<T>& e, const DevSpec<T>& sp, int k) {
  T cb = T(0);
#pragma unroll
  for (int c = 0; c < kF; ++c) cb = cb + e.B[c] * sp.m_sup[c][k];
  return cb;
}

template <typename T>
__device__ __forceinline__ void tenv_from_words(TEnv<T>& e, const T (&w)[kTStateWords]) {
#pragma unroll
  for (int k = 0; k < kF; ++k) {
    e.inv[k] = w[7 * k];
    e.unf_ind[k] = w[7 * k + 1];
    e.latI[k] = w[7 * k + 2];
#pragma unroll
    for (int j = 0; j < 4; ++j) e.U[k][j] = w[7 * k + 3 + j];
#pragma unroll
    for (int j = 0; j < kGI; ++j) e.info[k][j] = w[28 + kGI * k + j];
#pragma unroll
    for (int j = 0; j < kGS; ++j) e.ship[k][j] = w[40 + kGS * k + j];
    e.B[k] = w[56 + k];
    e.dem[k] = w[60 + k];
  }
}

template <typename T>
__device__ __forceinline__ void tenv_to_words(const TEnv<T>& e, T (&w)[kTStateWords]) {
#pragma unroll
  for (int k = 0; k < kF; ++k) {
    w[7 * k] = e.inv[k];
    w[7 * k + 1] = e.unf_ind[k];
    w[7 * k + 2] = e.latI[k];
#pragma unroll
    for (int j = 0; j < 4; ++j) w[7 * k + 3 + j] = e.U[k][j];
#pragma unroll
    for (int j = 0; j < kGI; ++j) w[28 + kGI * k + j] = e.info[k][j];
#pragma unroll
    for (int j = 0; j < kGS; ++j) w[40 + kGS * k + j] = e.ship[k][j];
    w[56 + k] = e.B[k];
    w[60 + k] = e.dem[k];
  }
}

// latest internal demand of every node from the slips that just reached the suppliers
template <typename T>
__device__ __forceinline__ void set_latest(TEnv<T>& e, const DevSpec<T>& sp, const T (&arrived)[kF]) {
#pragma unroll
  for (int j = 0; j < kF; ++j) {
    T v = T(0);
#pragma unroll
    for (int k = 0; k < kF; ++k) v = v + arrived[k] * sp.m_lat[j][k];
    e.latI[j] = v;                       // nodes without customers keep sum([]) = 0 (network_components.py:309)
  }
}

// Arc.reset / Node.reset + before_action(0)
template <typename T>
__device__ __forceinline__ void tenv_reset(TEnv<T>& e, const DevSpec<T>& sp, const T (&in)[kGInitWords], const TreeDem<T>& d0) {
  T arrived[kF];
#pragma unroll
  for (int k = 0; k < kF; ++k) {
    const int li = sp.li[k], ls = sp.ls[k];
    const bool real = k < sp.nf;
    T so[4], sh[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      so[j] = in[4 + 4 * k + j] * mask_of<T>(real && j < li);
      sh[j] = in[20 + 4 * k + j] * mask_of<T>(real && j < ls);
    }
    e.inv[k] = real ? in[k] : sp.big;
    T seq[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      T v = T(0);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        v = v + so[j] * mask_of<T>(li - 1 - i == j);
        v = v + sh[j] * mask_of<T>(i >= li && ls - 1 - (i - li) == j);
      }
      seq[i] = v;
    }
    e.U[k][0] = seq[0]; e.U[k][1] = seq[1]; e.U[k][2] = seq[2]; e.U[k][3] = seq[3] + seq[4];
    e.B[k] = so[0];
    arrived[k] = so[0];
#pragma unroll
    for (int j = 0; j < kGI; ++j) e.info[k][j] = so[j + 1];
#pragma unroll
    for (int j = 0; j < kGS; ++j) e.ship[k][j] = sh[j];
    e.unf_ind[k] = T(0);
    e.dem[k] = d0.v[k];
  }
  set_latest(e, sp, arrived);
}

template <typename T>
__device__ __forceinline__ void tenv_step(TEnv<T>& e, const DevSpec<T>& sp, const T (&q)[kF], const TreeDem<T>& d_next,
                                          bool terminal, StepOut& out) {
  // ---- receipts on every arc (supplynetwork.py:243-255)
  T U4[kF];
#pragma unroll
  for (int k = 0; k < kF; ++k) {
    T u[5] = {q[k], e.U[k][0], e.U[k][1], e.U[k][2], e.U[k][3]};
    const T arr = e.ship[k][0];
    e.inv[k] = e.inv[k] + arr;
    T rem = arr;
#pragma unroll
    for (int i = 4; i >= 0; --i) {
      const T f = tmin(u[i], rem);
      u[i] = u[i] - f;
      rem = rem - f;
    }
#pragma unroll
    for (int j = 0; j < kGS; ++j) e.ship[k][j] = (j + 1 < kGS) ? e.ship[k][j + 1] : T(0);
    e.U[k][0] = u[0]; e.U[k][1] = u[1]; e.U[k][2] = u[2]; e.U[k][3] = u[3];
    U4[k] = u[4];
  }
  // ---- fills, one round per customer rank (Arc.fill_orders in aggregate: min(inventory, backlog), :260-262)
#pragma unroll
  for (int r = 0; r < kF - 1; ++r) {
    if (r <= sp.max_rank) {                        // uniform over the grid
#pragma unroll
      for (int k = 0; k < kF; ++k) {
        T inv_s = sp.big * sp.m_ext[k];            // external supplier: unlimited
#pragma unroll
        for (int j = 0; j < kF; ++j) inv_s = inv_s + e.inv[j] * sp.m_sup[k][j];
        const T amt = tmin(inv_s, e.B[k]) * sp.m_rank[r][k];
#pragma unroll
        for (int j = 0; j < kF; ++j) e.inv[j] = e.inv[j] - amt * sp.m_sup[k][j];
        e.B[k] = e.B[k] - amt;
#pragma unroll
        for (int j = 0; j < kGS; ++j) e.ship[k][j] = e.ship[k][j] + amt * sp.m_ls[k][j];
      }
    }
  }
  // ---- independent demand (network_components.py:323-345); nodes without a stream have dem = unf_ind = 0
#pragma unroll
  for (int k = 0; k < kF; ++k) {
    const T tot = e.dem[k] + e.unf_ind[k];
    const T sold = tmax(T(0), tmin(e.inv[k], tot));
    e.inv[k] = e.inv[k] - sold;
    e.unf_ind[k] = tot - sold;
  }
  // ---- get_cost (supplynetwork.py:285-310)
  double ch = 0.0, cb = 0.0;
#pragma unroll
  for (int k = 0; k < kF; ++k) {
    const T unfk = customers_backlog(e, sp, k) + e.unf_ind[k];
    out.hold[k] = -(double)e.inv[k] * sp.h[0][k];
    out.back[k] = -(double)unfk * sp.b[0][k];
    ch += out.hold[k];
    cb += out.back[k];
  }
  out.reward = ch + cb;
#pragma unroll
  for (int k = 0; k < kF; ++k) e.dem[k] = d_next.v[k];                         // Node.update_demand
  if (!terminal) {
    T arrived[kF];
#pragma unroll
    for (int k = 0; k < kF; ++k) {
      arrived[k] = e.info[k][0] + q[k] * sp.m_li1[k];       // lead time 1: the head slot is empty, q arrives now
#pragma unroll
      for (int j = 0; j < kGI; ++j) {
        const T next = (j + 1 < kGI) ? e.info[k][j + 1] : T(0);
        e.info[k][j] = next + q[k] * sp.m_li[k][j];
      }
      e.B[k] = e.B[k] + arrived[k];
      e.U[k][3] = e.U[k][3] + U4[k];
    }
    set_latest(e, sp, arrived);
  }
}

}  // namespace sn
