// supplynet_generic.cuh — the same recurrence for serial chains with ARBITRARY per-arc lead times
// (first step: Arc(information_leadtime, shipment_leadtime) other than the beer game's
// 2,2,2,1 / 2,2,2,2).  The tuned `Env<T>` of supplynet_core.cuh hard-wires those; `GEnv<T>` keeps up to
// 3 in-flight order slips and 4 shipment slots per arc and takes the lead times from the spec at run time
// (slot positions are applied with 0/1 mask arithmetic over unrolled slots: registers only, no local memory).
//
// Supported: chains of 1..4 facilities (shorter chains pad the missing upstream nodes with phantoms, see
// DevSpec::nf), information and shipment lead times in 1..4 with info + shipment <= 5 per arc.  Beyond that
// the reference itself fails its on_order == sum(unreceived) assertion at reset (supplynetwork.py:159-160),
// because Arc.reset keeps only five pipeline slots (network_components.py:168).
//
// State word map of the generic path ([SN_GEN_STATE_WORDS = 57][ld]):
//   0..27   observation words exactly as in supplynet_core.cuh
//   28+3k+j info[k][j]   order slip of arc k reaching its supplier after j+1 more before_actions (j < Li-1)
//   40+4k+j ship[k][j]   shipment on arc k arriving after j+1 more sweeps (j < Ls)
//   56      B[3]
// Init words ([SN_GEN_INIT_WORDS = 36][ld]): inv[4]; so[k][j] at 4+4k+j; sh[k][j] at 20+4k+j (unused = 0).
#pragma once
#include "supplynet_core.cuh"

namespace sn {

constexpr int kGI = 3;                                   // live slips per arc after before_action (Li - 1 <= 3)
constexpr int kGS = 4;                                   // shipment slots per arc
constexpr int kGStateWords = 28 + kF * kGI + kF * kGS + 1;   // 57
constexpr int kGInitWords = kF + 4 * kF + 4 * kF;            // 36 (72 with eight nodes)

template <typename T>
struct GEnv {
  T inv[kF], lat[kF], U[kF][4], B[kF], info[kF][kGI], ship[kF][kGS], unf_ind;
  __device__ __forceinline__ T unfilled(int k, const DevSpec<T>&) const { return k == 0 ? lat[0] + unf_ind : B[k - 1]; }
  __device__ __forceinline__ T latest(int k) const { return lat[k]; }
  static __device__ __forceinline__ bool pre_agent(int k, const DevSpec<T>& sp) { return k < kF - 1 && k < sp.lo; }
  __device__ __forceinline__ uint32_t extra_qty_mask() const { return (uint32_t)unf_ind | (uint32_t)lat[0]; }
};

template <typename T>
__device__ __forceinline__ void genv_from_words(GEnv<T>& e, const T (&w)[kGStateWords]) {
#pragma unroll
  for (int k = 0; k < kF; ++k) {
    e.inv[k] = w[7 * k];
    if (k == 0) e.unf_ind = w[1]; else e.B[k - 1] = w[7 * k + 1];
    e.lat[k] = w[7 * k + 2];
#pragma unroll
    for (int j = 0; j < 4; ++j) e.U[k][j] = w[7 * k + 3 + j];
#pragma unroll
    for (int j = 0; j < kGI; ++j) e.info[k][j] = w[28 + kGI * k + j];
#pragma unroll
    for (int j = 0; j < kGS; ++j) e.ship[k][j] = w[40 + kGS * k + j];
  }
  e.B[3] = w[56];
}

template <typename T>
__device__ __forceinline__ void genv_to_words(const GEnv<T>& e, T (&w)[kGStateWords]) {
#pragma unroll
  for (int k = 0; k < kF; ++k) {
    w[7 * k] = e.inv[k];
    w[7 * k + 1] = (k == 0) ? e.unf_ind : e.B[k - 1];
    w[7 * k + 2] = e.lat[k];
#pragma unroll
    for (int j = 0; j < 4; ++j) w[7 * k + 3 + j] = e.U[k][j];
#pragma unroll
    for (int j = 0; j < kGI; ++j) w[28 + kGI * k + j] = e.info[k][j];
#pragma unroll
    for (int j = 0; j < kGS; ++j) w[40 + kGS * k + j] = e.ship[k][j];
  }
  w[56] = e.B[3];
}

// Run-time slot positions are applied with 0/1 mask arithmetic (x += m * v) instead of compare-selects
// over the slot arrays: nvcc turns select chains over array elements into dynamically indexed local-memory
// accesses (LDL/STL), which cost this path 3x (profiles/r1_notes.md).  Multiplying by an exact 0 or 1 and
// adding into a slot that is known to be empty is exact for integers and for finite floats.
template <typename T>
__device__ __forceinline__ T mask_of(bool c) { return c ? T(1) : T(0); }

// Arc.reset / Node.reset + before_action(0) for run-time lead times (network_components.py:138-168).
template <typename T>
__device__ __forceinline__ void genv_reset(GEnv<T>& e, const DevSpec<T>& sp, const T (&in)[kGInitWords], T d0) {
#pragma unroll
  for (int k = 0; k < kF; ++k) {
    const int li = sp.li[k], ls = sp.ls[k];
    const bool real = k < sp.nf;                    // phantom nodes of a shorter chain: empty pipelines, stock `big`
    T so[4], sh[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      so[j] = in[4 + 4 * k + j] * mask_of<T>(real && j < li);
      sh[j] = in[20 + 4 * k + j] * mask_of<T>(real && j < ls);
    }
    e.inv[k] = real ? in[k] : sp.big;
    // ([0]*4 + shipments + sales_orders)[::-1][:5]: newest slip first, then shipments latest first
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
    // before_action(0): the list loses its last slot into its neighbour; so[0] reaches the supplier
    e.U[k][0] = seq[0]; e.U[k][1] = seq[1]; e.U[k][2] = seq[2]; e.U[k][3] = seq[3] + seq[4];
    e.B[k] = so[0];
    if (k < 3) e.lat[k + 1] = so[0];
#pragma unroll
    for (int j = 0; j < kGI; ++j) e.info[k][j] = so[j + 1];        // already 0 beyond li
#pragma unroll
    for (int j = 0; j < kGS; ++j) e.ship[k][j] = sh[j];
  }
  e.unf_ind = T(0);
  e.lat[0] = d0;
}

// One period; same algebra as env_step with run-time pipeline lengths.
template <typename T>
__device__ __forceinline__ void genv_step(GEnv<T>& e, const DevSpec<T>& sp, int item, const T (&q)[kF], T d_next,
                                          bool terminal, StepOut& out) {
  T U4[kF];
#pragma unroll
  for (int k = kF - 1; k >= 0; --k) {
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
    T s;
    if (k == kF - 1) {
      s = e.B[k];
    } else {
      s = tmin(e.inv[k + 1], e.B[k]);
      e.inv[k + 1] = e.inv[k + 1] - s;
    }
    // advance the shipments and append Shipment(s, shipment_leadtime): slot ls-1 is empty after the shift
#pragma unroll
    for (int j = 0; j < kGS; ++j) {
      const T next = (j + 1 < kGS) ? e.ship[k][j + 1] : T(0);
      e.ship[k][j] = next + s * sp.m_ls[k][j];
    }
    e.B[k] = e.B[k] - s;
    e.U[k][0] = u[0]; e.U[k][1] = u[1]; e.U[k][2] = u[2]; e.U[k][3] = u[3];
    U4[k] = u[4];
  }
  const T tot = e.lat[0] + e.unf_ind;
  const T sold = tmax(T(0), tmin(e.inv[0], tot));
  e.inv[0] = e.inv[0] - sold;
  e.unf_ind = tot - sold;

  double ch = 0.0, cb = 0.0;
#pragma unroll
  for (int k = 0; k < kF; ++k) {
    const T unfk = (k == 0) ? e.unf_ind : e.B[k - 1];
    out.hold[k] = __dmul_rn(-(double)e.inv[k], sp.h[item][k]);        // item: cost vector of this chain (0 for single-item batches)
    out.back[k] = __dmul_rn(-(double)unfk, sp.b[item][k]);
    ch = __dadd_rn(ch, out.hold[k]);
    cb = __dadd_rn(cb, out.back[k]);
  }
  out.reward = __dadd_rn(ch, cb);

  e.lat[0] = d_next;
  if (!terminal) {
#pragma unroll
    for (int k = 0; k < kF; ++k) {
      // Order(q, 0, info_leadtime): with lead time 1 this period's order is what arrives now (head is 0 then)
      const T arrived = e.info[k][0] + q[k] * sp.m_li1[k];
#pragma unroll
      for (int j = 0; j < kGI; ++j) {
        const T next = (j + 1 < kGI) ? e.info[k][j + 1] : T(0);
        e.info[k][j] = next + q[k] * sp.m_li[k][j];
      }
      if (k < 3) e.lat[k + 1] = arrived;
      e.B[k] = e.B[k] + arrived;
      e.U[k][3] = e.U[k][3] + U4[k];
    }
  }
}

}  // namespace sn
