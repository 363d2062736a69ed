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
// SPECIALISED BUILDS (-DSN_STATIC_TOPO, multi_echelon_drl_b200/specialize.py): the supplier table, the customer ranks,
// the latest-demand sources and the lead times of ONE network are baked in as constants (SN_ST_*), every `pick` below
// folds at compile time, and what is left is the straight-line code of that network - the 0/1-multiplier arithmetic
// of the run-time path (861 instructions per warp-period against 342 for the tuned chain) disappears.  Same state
// layout, same results bit for bit; the entry points of such a library refuse any other network.
#pragma once
#include "supplynet_generic.cuh"

namespace sn {

// state / init word maps.  Four nodes: the map above (64 words).  Eight nodes (specialised builds only): the map of the
// lane-per-node kernels, 16 words per node (supplynet_net.cuh), so that reset / observe of the general library and
// the specialised step / rollout share one state buffer.  Init words: inv[kF]; so[k][0..3] at kF+4k; sh[k][0..3]
// at 5kF+4k in both (36 / 72 words).
constexpr int kTStateWords = kF == 4 ? 64 : 16 * kF;
__host__ __device__ constexpr int tw_inv(int k) { return kF == 4 ? 7 * k : 16 * k; }
__host__ __device__ constexpr int tw_unf(int k) { return tw_inv(k) + 1; }
__host__ __device__ constexpr int tw_lat(int k) { return tw_inv(k) + 2; }
__host__ __device__ constexpr int tw_U(int k, int j) { return tw_inv(k) + 3 + j; }
__host__ __device__ constexpr int tw_info(int k, int j) { return kF == 4 ? 28 + kGI * k + j : 16 * k + 7 + j; }
__host__ __device__ constexpr int tw_ship(int k, int j) { return kF == 4 ? 40 + kGS * k + j : 16 * k + 10 + j; }
__host__ __device__ constexpr int tw_B(int k) { return kF == 4 ? 56 + k : 16 * k + 14; }
__host__ __device__ constexpr int tw_dem(int k) { return kF == 4 ? 60 + k : 16 * k + 15; }

#ifdef SN_STATIC_TOPO
constexpr bool kStaticTopo = true;
__host__ __device__ constexpr int st8(int k, int a, int b, int c, int d, int e, int f, int g, int h) {
  return k == 0 ? a : k == 1 ? b : k == 2 ? c : k == 3 ? d : k == 4 ? e : k == 5 ? f : k == 6 ? g : h;
}
#define SN_ST8(NAME) st8(k, SN_ST_##NAME##0, SN_ST_##NAME##1, SN_ST_##NAME##2, SN_ST_##NAME##3, SN_ST_##NAME##4, \
                         SN_ST_##NAME##5, SN_ST_##NAME##6, SN_ST_##NAME##7)
__host__ __device__ constexpr int st_nf() { return SN_ST_NF; }
__host__ __device__ constexpr int st_sup(int k) { return SN_ST8(SUP); }
__host__ __device__ constexpr int st_rank(int k) { return SN_ST8(RANK); }
__host__ __device__ constexpr int st_lat(int k) { return SN_ST8(LAT); }
__host__ __device__ constexpr int st_li(int k) { return SN_ST8(LI); }
__host__ __device__ constexpr int st_ls(int k) { return SN_ST8(LS); }
__host__ __device__ constexpr int st_max_rank() { return SN_ST_MAX_RANK; }
#else
constexpr bool kStaticTopo = false;
__host__ __device__ constexpr int st_nf() { return kF; }
__host__ __device__ constexpr int st_sup(int) { return 0; }
__host__ __device__ constexpr int st_rank(int) { return 0; }
__host__ __device__ constexpr int st_lat(int) { return 0; }
__host__ __device__ constexpr int st_li(int) { return 0; }
__host__ __device__ constexpr int st_ls(int) { return 0; }
__host__ __device__ constexpr int st_max_rank() { return kF; }
#endif

// v (+|-)= x where the table says so: a 0/1 multiplier from the constant bank at run time (no branch, no register
// select), a compile-time condition in specialised builds (`c` is a constant after unrolling: the statement either
// stays as a plain add or vanishes)
template <typename T>
__device__ __forceinline__ void pick_add(T& v, T x, T m, bool c) {
  if (kStaticTopo) { if (c) v = v + x; } else { v = v + x * m; }
}
template <typename T>
__device__ __forceinline__ void pick_sub(T& v, T x, T m, bool c) {
  if (kStaticTopo) { if (c) v = v - x; } else { v = v - x * m; }
}

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
    for (int c = 0; c < kF; ++c) pick_add(cb, B[c], sp.m_sup[c][k], st_sup(c) == k);
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
  for (int c = 0; c < kF; ++c) pick_add(cb, e.B[c], sp.m_sup[c][k], st_sup(c) == k);
  return cb;
}

template <typename T>
__device__ __forceinline__ void tenv_from_words(TEnv<T>& e, const T (&w)[kTStateWords]) {
#pragma unroll
  for (int k = 0; k < kF; ++k) {
    e.inv[k] = w[tw_inv(k)];
    e.unf_ind[k] = w[tw_unf(k)];
    e.latI[k] = w[tw_lat(k)];
#pragma unroll
    for (int j = 0; j < 4; ++j) e.U[k][j] = w[tw_U(k, j)];
#pragma unroll
    for (int j = 0; j < kGI; ++j) e.info[k][j] = w[tw_info(k, j)];
#pragma unroll
    for (int j = 0; j < kGS; ++j) e.ship[k][j] = w[tw_ship(k, j)];
    e.B[k] = w[tw_B(k)];
    e.dem[k] = w[tw_dem(k)];
  }
}

template <typename T>
__device__ __forceinline__ void tenv_to_words(const TEnv<T>& e, T (&w)[kTStateWords]) {
#pragma unroll
  for (int k = 0; k < kF; ++k) {
    w[tw_inv(k)] = e.inv[k];
    w[tw_unf(k)] = e.unf_ind[k];
    w[tw_lat(k)] = e.latI[k];
#pragma unroll
    for (int j = 0; j < 4; ++j) w[tw_U(k, j)] = e.U[k][j];
#pragma unroll
    for (int j = 0; j < kGI; ++j) w[tw_info(k, j)] = e.info[k][j];
#pragma unroll
    for (int j = 0; j < kGS; ++j) w[tw_ship(k, j)] = e.ship[k][j];
    w[tw_B(k)] = e.B[k];
    w[tw_dem(k)] = e.dem[k];
  }
}

// Specialised builds: pipeline slots beyond an arc's lead time never hold anything (reset leaves them 0, the step only
// ever adds into slot lead time - 1 and shifts down).  Saying so after every load lets the compiler drop them from the
// loop: up to 18 of the 64 state registers for the beer-game-like lead times.
template <typename T>
__device__ __forceinline__ void tenv_static_zeros(TEnv<T>& e) {
  if (kStaticTopo) {
#pragma unroll
    for (int k = 0; k < kF; ++k) {
#pragma unroll
      for (int j = 0; j < kGI; ++j)
        if (k >= st_nf() || j >= st_li(k) - 1) e.info[k][j] = T(0);
#pragma unroll
      for (int j = 0; j < kGS; ++j)
        if (k >= st_nf() || j >= st_ls(k)) e.ship[k][j] = T(0);
    }
  }
}

// latest internal demand of every node from the slips that just reached the suppliers
template <typename T>
__device__ __forceinline__ void set_latest(TEnv<T>& e, const DevSpec<T>& sp, const T (&arrived)[kF]) {
#pragma unroll
  for (int j = 0; j < kF; ++j) {
    T v = T(0);
#pragma unroll
    for (int k = 0; k < kF; ++k) pick_add(v, arrived[k], sp.m_lat[j][k], st_lat(j) == k);
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
      so[j] = in[kF + 4 * k + j] * mask_of<T>(real && j < li);
      sh[j] = in[5 * kF + 4 * k + j] * mask_of<T>(real && j < ls);
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
    if (kStaticTopo && k >= st_nf()) { U4[k] = T(0); continue; }     // phantom node: nothing moves
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
    if (kStaticTopo ? r <= st_max_rank() : r <= sp.max_rank) {                   // uniform over the grid
#pragma unroll
      for (int k = 0; k < kF; ++k) {
        if (kStaticTopo && (k >= st_nf() || st_rank(k) != r)) continue;          // node k is served in round rank[k] only
        T inv_s = kStaticTopo ? (st_sup(k) < 0 ? sp.big : T(0)) : sp.big * sp.m_ext[k];      // external supplier: unlimited
#pragma unroll
        for (int j = 0; j < kF; ++j) pick_add(inv_s, e.inv[j], sp.m_sup[k][j], st_sup(k) == j);
        const T amt = kStaticTopo ? tmin(inv_s, e.B[k]) : tmin(inv_s, e.B[k]) * sp.m_rank[r][k];
#pragma unroll
        for (int j = 0; j < kF; ++j) pick_sub(e.inv[j], amt, sp.m_sup[k][j], st_sup(k) == j);
        e.B[k] = e.B[k] - amt;
#pragma unroll
        for (int j = 0; j < kGS; ++j) pick_add(e.ship[k][j], amt, sp.m_ls[k][j], st_ls(k) - 1 == j);
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
    out.hold[k] = __dmul_rn(-(double)e.inv[k], sp.h[0][k]);
    out.back[k] = __dmul_rn(-(double)unfk, sp.b[0][k]);
    ch = __dadd_rn(ch, out.hold[k]);
    cb = __dadd_rn(cb, out.back[k]);
  }
  out.reward = __dadd_rn(ch, cb);
#pragma unroll
  for (int k = 0; k < kF; ++k) e.dem[k] = d_next.v[k];                         // Node.update_demand
  if (!terminal) {
    T arrived[kF];
#pragma unroll
    for (int k = 0; k < kF; ++k) {
      arrived[k] = e.info[k][0];                             // lead time 1: the head slot is empty, q arrives now
      pick_add(arrived[k], q[k], sp.m_li1[k], st_li(k) == 1);
#pragma unroll
      for (int j = 0; j < kGI; ++j) {
        T next = (j + 1 < kGI) ? e.info[k][j + 1] : T(0);
        pick_add(next, q[k], sp.m_li[k][j], st_li(k) - 2 == j);
        e.info[k][j] = next;
      }
      e.B[k] = e.B[k] + arrived[k];
      e.U[k][3] = e.U[k][3] + U4[k];
    }
    set_latest(e, sp, arrived);
  }
}

}  // namespace sn
