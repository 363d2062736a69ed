/*
 * supplynet.h — C ABI of the B200-native batched supply-network (beer game) simulator.
 *
 * The reference (qihuazhong/multi-echelon-drl) is pure Python and has no FFI; this header
 * is the boundary a maintainer would bind with ctypes.  Each entry
 * point names the reference interface it replaces (file:line into the reference tree).
 *
 * Conventions
 *   - Every pointer marked "device" is caller-owned CUDA device memory (e.g. a torch
 *     tensor's data_ptr()).  The library never allocates or frees device memory and keeps
 *     no state between calls: all simulator state lives in the caller's `state` buffer.
 *   - All launches are asynchronous on the caller's stream (`stream` is a cudaStream_t
 *     passed as void*; NULL = legacy default stream).
 *   - Structure-of-arrays: a "row" is one quantity for all envs, `ld` elements apart
 *     (ld >= n_envs, ld % 4 == 0, base pointers 16-byte aligned).
 *   - `dtype` selects the arithmetic type of quantities: SN_I32 (bit-exact integer
 *     rollouts), SN_F32 / SN_F64 (fractional orders, e.g. TD3 actions).
 *   - Return value: SN_OK (0) or a negative SN_ERR_* code; sn_last_error() gives text.
 *   - There is no CPU fallback: without a CUDA device every compute call fails.
 *   - Environment switches, read per call, for A/B runs and tests only: SN_STEP_KERNEL=plain|tiled forces one of the
 *     two implementations of the period step (default: by batch size, see sn_step); SN_NO_COOP=1 launches the
 *     one-kernel VecNormalize mode without the cooperative attribute (experiments).
 */
#ifndef SUPPLYNET_H
#define SUPPLYNET_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SN_ABI_VERSION 11

#define SN_MAX_FACILITIES 8     /* size of the per-facility arrays; chains and the register-resident kernels use up to 4 */
#define SN_MAX_CHAIN 4          /* chains, and trees on the register-resident kernels */
#define SN_MAX_ITEMS 4
#define SN_OBS_PER_FACILITY 7   /* supplynetwork.py:151-169 */
#define SN_STATE_WORDS 40       /* 28 observable + 12 hidden words per env */
#define SN_INIT_WORDS 19        /* inv[4], sales orders (2,2,2,1), shipments (2,2,2,2) */
#define SN_GEN_STATE_WORDS 57   /* chains with other lead times: 28 + 4x3 order slips + 4x4 shipments + 1 */
#define SN_GEN_INIT_WORDS 36    /* inv[4]; so[k][0..3] at 4+4k; sh[k][0..3] at 20+4k (unused slots 0)      */
#define SN_TREE_STATE_WORDS 64  /* distribution trees: 28 + 12 + 16 + backlog[4] + current external demand[4] */
#define SN_NET_STATE_WORDS 128  /* trees of 5..8 nodes: 16 words per node (inv, unf_ind, latest, U[4], info[3], ship[4], backlog, demand) */
#define SN_NET_INIT_WORDS 72    /* trees of 5..8 nodes: inv[8]; so[k][0..3] at 8+4k; sh[k][0..3] at 40+4k */
#define SN_STATS_WORDS 16
#define SN_QTY_LIMIT_LOG2 26    /* SN_I32 is exact while orders, stocks and backlogs stay below 2^26 (see stats[13]) */

/* Device control words of sn_step_ex (int32[SN_CTL_WORDS], caller-owned, zero-initialised):
 * the period lives in device memory so that ONE captured CUDA graph serves every period of an episode. */
#define SN_CTL_WORDS 8
#define SN_CTL_PERIOD 0         /* period to play; read by the launch and advanced by it (set to 0 to start an episode) */
#define SN_CTL_TICKET 1         /* internal (block ticket); leave at 0                                               */
#define SN_CTL_FLAGS 2          /* sticky SN_FLAG_* bits                                                             */
#define SN_CTL_SEQ 4            /* internal (launch sequence number of the one-kernel VecNormalize mode); never reset */
#define SN_VN_ONEPASS_SCRATCH_DOUBLES (256 * 128)  /* vn_scratch of that mode: per block 64 entries of 16 bytes         */
#define SN_FLAG_STEP_AFTER_TERMINAL 1   /* a launch found period >= max_episode_steps and did nothing (envs.py:88-89) */
#define SN_FLAG_QTY_RANGE 2             /* SN_I32: a quantity reached 2^SN_QTY_LIMIT_LOG2, results may have wrapped    */

enum { SN_I32 = 0, SN_F32 = 1, SN_F64 = 2 };
enum { SN_RULE_A = 0, SN_RULE_D_PLUS_A = 1 };
enum { SN_TOPO_CHAIN = 0, SN_TOPO_TREE = 1 };
#define SN_EXTERNAL (-1)
enum {
  SN_POLICY_ACTIONS = 0,      /* actions streamed from memory, [T][N][n_agents]            */
  SN_POLICY_BASE_STOCK = 1    /* utils/heuristics.py base-stock evaluated in-kernel per env */
};

enum {
  SN_OK = 0,
  SN_ERR_INVALID = -1,      /* bad argument (maps to ValueError)                          */
  SN_ERR_UNSUPPORTED = -2,  /* topology / lead times outside the serial beer-game chain   */
  SN_ERR_TERMINAL = -3,     /* step after the episode ended (envs.py:88-89 ValueError)     */
  SN_ERR_CUDA = -4,         /* CUDA runtime error                                          */
  SN_ERR_TYPE = -5          /* non-integral constant with SN_I32 (maps to TypeError)       */
};

/* Scenario description: the constants hard-coded in the reference's factories
 * (envs.py:163-286 "normal"/basic, :289-414 "uniform"/complex, :501-596 classic) plus the
 * ordering rule of utils/wrappers.py:6-31.  Node k: 0 retailer .. 3 manufacturer; arc k is the
 * arc whose customer is node k. */
typedef struct sn_spec {
  int32_t n_facilities;                       /* chains 1..4: the first n of retailer..manufacturer, arc n-1 buys from the
                                               * external supplier; trees 1..8 internal nodes */
  int32_t info_leadtime[SN_MAX_FACILITIES];   /* Arc.information_leadtime, 1..4; beer game 2,2,2,1 */
  int32_t ship_leadtime[SN_MAX_FACILITIES];   /* Arc.shipment_leadtime, 1..4, info+ship <= 5; 2,2,2,2 */
  int32_t n_agents;                           /* 1..n_facilities                              */
  int32_t agents[SN_MAX_FACILITIES];          /* node index per action column, caller order   */
                                              /* (supplynetwork.py:217); contiguous set       */
  int32_t global_observable;                  /* envs.py:72-75                                */
  int32_t max_episode_steps;                  /* envs.py:98                                   */
  int32_t rule;                               /* SN_RULE_*                                    */
  double holding_cost[SN_MAX_FACILITIES];     /* Node.unit_holding_cost                       */
  double backorder_cost[SN_MAX_FACILITIES];   /* Node.unit_backorder_cost                     */
  double coplayer_target[SN_MAX_FACILITIES];  /* BaseStockPolicy target of non-agent nodes    */
  double coplayer_lb, coplayer_ub;            /* utils/heuristics.py:16 (0, 16)               */
  double dpa_offset, dpa_lb, dpa_ub;          /* utils/wrappers.py:6                          */
  /* Multi-item networks (network_configs/multi_item.yaml: per-item cost lists).  Items are independent
   * copies of the chain: with n_items = I > 1 the batch holds n_envs*I chains, chain c = env*I + item
   * (item-minor), so the row-major observation [n_chains][7*F'] IS the concatenation over items
   * [n_envs][I*7*F'] and likewise for actions; rewards are per chain.  Item 0 uses holding_cost /
   * backorder_cost above, item i >= 1 uses item_*_cost[i-1].  n_items = 0 is read as 1.  Serial chains only (any
   * supported lead times and chain length); trees are single-item. */
  int32_t n_items;
  int32_t topology;                           /* SN_TOPO_CHAIN (0): the fields below are ignored; SN_TOPO_TREE */
  double item_holding_cost[SN_MAX_ITEMS - 1][SN_MAX_FACILITIES];
  double item_backorder_cost[SN_MAX_ITEMS - 1][SN_MAX_FACILITIES];
  /* Distribution trees (SN_TOPO_TREE): every internal node k buys from ONE supplier and may serve several
   * customers and / or its own external demand.  This is what SupplyNetwork does for such node/arc lists
   * (supplynetwork.py:22-56, 182-268): node k is the k-th internal node of the node list (observation and cost
   * order); its upstream arc carries info_leadtime[k] / ship_leadtime[k].  The demand tables then hold one row
   * per demand source and period, sources in node order: demand [T+3][n_sources][ld], demand0 / demand_next
   * [n_sources][ld].  Multi-supplier (assembly) nodes are not supported: the reference's own state keys collide
   * there (supplynetwork.py:165-167).
   * Trees of up to four nodes run on one-thread-per-env kernels (SN_TREE_STATE_WORDS); trees of five to eight nodes on
   * kernels that give every node a lane of its own, eight lanes per env (SN_NET_STATE_WORDS / SN_NET_INIT_WORDS,
   * observation 7 floats per node, up to 8 action columns, `cost` [16][ld]: holding[0..7], backorder[8..15]; the
   * per-facility sums [2..9] of the statistics vector stay zero there). */
  int32_t supplier[SN_MAX_FACILITIES];        /* supplier node of node k, SN_EXTERNAL = the external supplier   */
  int32_t demand_source[SN_MAX_FACILITIES];   /* Node.is_demand_source                                           */
  int32_t order_pos[SN_MAX_FACILITIES];       /* position of node k among the internal nodes in the order sequence
                                               * (utils/graph.py:5-32): customers before suppliers; the agents
                                               * must be contiguous in it (supplynetwork.py:201,231)             */
  int32_t customer_rank[SN_MAX_FACILITIES];   /* rank of node k among its supplier's customers in customers_dict
                                               * order (= arc-list order, supplynetwork.py:22-25): the supplier
                                               * serves rank 0 first (Arc.fill_orders per arc, :260-262)         */
} sn_spec;

/* Base-stock policy parameters (utils/heuristics.py:13-33), used by sn_heuristic and by
 * sn_rollout with SN_POLICY_BASE_STOCK.  One target per action column. */
typedef struct sn_policy {
  double target[SN_MAX_FACILITIES];
  double lb, ub;
  int32_t rule;              /* SN_RULE_A: target-IP; SN_RULE_D_PLUS_A: target-IP-latest_demand */
  int32_t row0_broadcast;    /* 1 = reproduce the reference's batch quirk (row 0 for all envs)   */
} sn_policy;

int32_t sn_abi_version(void);
const char* sn_last_error(void);

/* Number of observation columns for this spec: 7 * (4 if global_observable else n_agents)
 * (envs.py:402-405).  Negative on invalid spec. */
int32_t sn_obs_dim(const sn_spec* spec);

/* Validates a spec for `dtype` (replaces the implicit checks of SupplyNetwork.__init__,
 * supplynetwork.py:22-56; rejects non-contiguous agent sets that crash the reference at :199). */
int32_t sn_spec_validate(const sn_spec* spec, int32_t dtype);

/* Rows of the `state` / `init` buffers for this spec: SN_STATE_WORDS / SN_INIT_WORDS for the beer game's
 * lead times (tuned kernels), SN_GEN_STATE_WORDS / SN_GEN_INIT_WORDS for any other supported lead times
 * (generic kernels, network_components.py:108-168 with Arc(information_leadtime, shipment_leadtime)),
 * SN_TREE_STATE_WORDS / SN_GEN_INIT_WORDS for SN_TOPO_TREE. */
int32_t sn_state_words(const sn_spec* spec);
int32_t sn_init_words(const sn_spec* spec);

/* reset — replaces InventoryManagementEnvMultiPlayer.reset (envs.py:64-80):
 * SupplyNetwork.reset (supplynetwork.py:94-105; Arc.reset network_components.py:138-168,
 * Node.reset :293-319) followed by before_action(0) (supplynetwork.py:182-210).
 *   init      device, dtype [sn_init_words(spec)][ld]: inv[0..3]; sales orders so[k][j] for k=0..3
 *             (j < info_leadtime[k], j = periods-1 until the supplier sees it); shipments
 *             sh[k][j] for k=0..3 (j = periods-1 until arrival)
 *   demand0   device, dtype [ld]: demand of period 0 (Node.reset :317-319)
 *   state     device, dtype [sn_state_words(spec)][ld] (out)
 *   obs       device, float [n_envs][obs_dim] row-major (out, may be NULL) */
int32_t sn_reset(const sn_spec* spec, int32_t dtype, int64_t n_envs, int64_t ld,
                 const void* init, const void* demand0, void* state, float* obs, void* stream);

/* step — replaces InventoryManagementEnvMultiPlayer.step (envs.py:82-113) for all envs in
 * lockstep, including the d+a action rule (utils/wrappers.py:23-27) when spec->rule says so.
 *   period        the period being played (0-based); the step is terminal when
 *                 period + 1 >= max_episode_steps, in which case `obs` receives the terminal
 *                 observation (no before_action; envs.py:98-101) and `state` is left at an
 *                 unspecified post-terminal value: call sn_reset next.
 *   actions       device, dtype [n_envs][n_agents] row-major (what VecEnv.step receives)
 *   demand_next   device, dtype [ld]: demand of period+1 (Node.update_demand,
 *                 network_components.py:347-348)
 *   reward        device, float [n_envs] (out): SupplyNetwork.get_cost (supplynetwork.py:285-310),
 *                 summed in double in the reference's order, rounded once to float
 *   reward64      device, double [n_envs] (out, may be NULL): the unrounded value
 *   cost          device, double [8][ld] (out, may be NULL): per-facility holding[0..3] and
 *                 backorder[4..7] cost of this period (the cost-history entries :303-305)
 *   ep_return     device, double [n_envs] (in/out, may be NULL): += reward (Monitor's "r")
 *   stats         device, double [SN_STATS_WORDS] (in/out, may be NULL): += totals over envs:
 *                 [0] env-steps, [1] sum reward, [2..5] sum holding cost per facility,
 *                 [6..9] sum backorder cost per facility (warp/block reduced, one atomic per block);
 *                 when the step ends the episode and ep_return is given also [10] episodes,
 *                 [11] sum of episode returns, [12] sum of squared episode returns;
 *                 [13] SN_I32 only: env-steps in which an action magnitude, order, inventory or backlog reached
 *                 2^SN_QTY_LIMIT_LOG2 (the reference computes with unbounded Python ints; beyond that bound int32
 *                 intermediates may wrap, so a non-zero count voids the bit-exactness claim for that run); sn_rollout /
 *                 sn_episode count envs instead of env-steps and do not watch distribution trees of up to four nodes
 *                 (their rollout kernel has no register to spare: step such a network with sn_step to have it watched)
 * Returns SN_ERR_TERMINAL if period >= max_episode_steps. */
int32_t sn_step(const sn_spec* spec, int32_t dtype, int64_t n_envs, int64_t ld, int32_t period,
                void* state, const void* actions, const void* demand_next,
                float* obs, float* reward, double* reward64, double* cost, double* ep_return,
                double* stats, void* stream);

/* step, extended form — the same period step with the optional extras a device-resident training loop needs.
 * sn_step(...) == sn_step_ex with ctl = env_reward = vn_* = NULL.
 *   demand       ctl == NULL: the [ld] row(s) of period+1 exactly as sn_step's demand_next;
 *                ctl != NULL: the whole table [max_episode_steps+3][...] as in sn_rollout (the launch picks the row)
 *   period       host-side period; ignored when ctl != NULL
 *   ctl          device int32[SN_CTL_WORDS] or NULL.  The launch plays period ctl[SN_CTL_PERIOD] and advances it
 *                (envs.py:96), so identical launches (a replayed CUDA graph) walk through an episode; a launch that
 *                finds the episode over does nothing and raises SN_FLAG_STEP_AFTER_TERMINAL in ctl[SN_CTL_FLAGS]
 *   env_reward   device float [n_envs / n_items] (out, may be NULL; n_items 2 or 4): multi-item batches, reward of an
 *                env = sum of its items' chain rewards in item order, in double, rounded once
 *   vn_*         fused VecNormalize statistics (train_td3.py:100-101; needs ctl, beer-game lead times, n_items <= 1,
 *                obs != NULL): RunningMeanStd.update(obs) into vn_obs_rms (mean[od], var[od], count) IN PLACE and, if
 *                vn_returns != NULL, returns <- returns*vn_gamma + reward and RunningMeanStd.update(returns) into
 *                vn_ret_rms (mean, var, count); vn_scratch is 2*obs_dim+2 doubles, zero before the first launch (the
 *                launch leaves it zero).  Follow with sn_vecnorm_apply.  Not for the terminal step of an episode:
 *                SB3 feeds the statistics the observation of the NEXT episode there (DummyVecEnv auto-reset).
 *   vn_nobs      one-kernel mode (n_envs <= SMs x 128, i.e. one tile per SM; float32/int32 quantities): the same
 *                launch also writes the normalised observation [n_envs][obs_dim] (and vn_nrew [n_envs], if given):
 *                clip((obs-mean)/sqrt(var+vn_epsilon), +-vn_clip_obs), clip(reward/sqrt(ret_var+eps), +-vn_clip_reward),
 *                with the statistics AFTER this step's update when vn_scratch is given (then vn_scratch is
 *                SN_VN_ONEPASS_SCRATCH_DOUBLES doubles, zero before the first launch, and the launch is cooperative:
 *                every block publishes its tile's moments as 16-byte entries that carry value and launch tag together,
 *                reads everybody else's, and adds the blocks up in block order - one trip through the L2, and the
 *                statistics are reproducible bit for bit, unlike atomic accumulation; the buffer belongs to ONE ctl /
 *                VecNormalize pair: tags are the ctl's launch sequence numbers), else with vn_obs_rms / vn_ret_rms
 *                as they are (evaluation).  Replaces the sn_vecnorm_apply launch. */
typedef struct sn_step_io {
  void* state;
  const void* actions;
  const void* demand;
  int32_t period;
  int32_t reserved;
  int32_t* ctl;
  float* obs;
  float* reward;
  float* env_reward;
  double* reward64;
  double* cost;
  double* ep_return;
  double* stats;
  double* vn_scratch;
  double* vn_obs_rms;
  double* vn_ret_rms;
  double* vn_returns;
  double vn_gamma;
  float* vn_nobs;
  float* vn_nrew;
  float vn_clip_obs;
  float vn_clip_reward;
  double vn_epsilon;
} sn_step_io;
int32_t sn_step_ex(const sn_spec* spec, int32_t dtype, int64_t n_envs, int64_t ld, const sn_step_io* io, void* stream);

/* observation of the current state — SupplyNetwork.get_state (supplynetwork.py:114-171)
 * flattened as envs.py:103-110.  obs: device float [n_envs][obs_dim]. */
int32_t sn_observe(const sn_spec* spec, int32_t dtype, int64_t n_envs, int64_t ld,
                   const void* state, float* obs, void* stream);

/* heuristic — replaces BaseStockPolicy.get_order_quantity (utils/heuristics.py:35-75), the
 * ndarray path HgeTD3.predict calls (hge.py:157-167), evaluated per env.
 *   obs      device, float [n_envs][7*n_cols] row-major (un-normalised observations)
 *   actions  device, float [n_envs][n_cols] (out) */
int32_t sn_heuristic(const sn_policy* policy, int32_t n_cols, int64_t n_envs,
                     const float* obs, float* actions, void* stream);

/* rollout — n_steps consecutive periods fused in one launch with the per-env state held
 * on-chip (envs.py:82-113 iterated; policy = streamed actions or in-kernel base stock).
 *   period0       first period to play; period0 + n_steps <= max_episode_steps
 *   demand        device, dtype [max_episode_steps+3][ld] (utils/demands.py:21 "size+3"); row t is
 *                 the demand of period t
 *   actions       device, dtype [n_steps][n_envs][n_agents] (SN_POLICY_ACTIONS) or NULL
 *   policy        base-stock parameters (SN_POLICY_BASE_STOCK) or NULL
 *   traj_obs      device, float [n_steps][n_envs][obs_dim] (out, may be NULL)
 *   traj_reward   device, float [n_steps][n_envs] (out, may be NULL)
 *   traj_actions  device, float [n_steps][n_envs][n_agents] (out, may be NULL): the actions taken
 *   obs           device, float [n_envs][obs_dim] (out, may be NULL): observation after the last step
 *   ep_return     device, double [n_envs] (in/out, may be NULL)
 *   stats         device, double [SN_STATS_WORDS] (in/out, may be NULL), as in sn_step */
int32_t sn_rollout(const sn_spec* spec, int32_t dtype, int64_t n_envs, int64_t ld, int32_t period0,
                   int32_t n_steps, void* state, const void* demand, int32_t policy_kind,
                   const void* actions, const sn_policy* policy, float* traj_obs, float* traj_reward,
                   float* traj_actions, float* obs, double* ep_return, double* stats, void* stream);

/* episode — sn_reset + sn_rollout from period 0 fused in ONE launch: the env is reset from `init`
 * inside the kernel (envs.py:64-80) and stepped n_steps periods (envs.py:82-113) without the state
 * ever living in HBM.  This is the form used for fixed-action parity rollouts and for base-stock /
 * HGE heuristic rollouts (utils/heuristics.py) over whole episodes.
 *   init, demand            as in sn_reset / sn_rollout
 *   obs0                    device, float [n_envs][obs_dim] (out, may be NULL): what reset() returns
 *   state                   device, dtype [SN_STATE_WORDS][ld] (out, may be NULL): state after n_steps
 *   ep_return               device, double [n_envs] (out, may be NULL): OVERWRITTEN with the sum of rewards
 *   other arguments         as in sn_rollout */
int32_t sn_episode(const sn_spec* spec, int32_t dtype, int64_t n_envs, int64_t ld, int32_t n_steps,
                   const void* init, const void* demand, int32_t policy_kind, const void* actions,
                   const sn_policy* policy, float* obs0, float* traj_obs, float* traj_reward,
                   float* traj_actions, float* obs, void* state, double* ep_return, double* stats,
                   void* stream);

/* ---- device-side VecNormalize (stable_baselines3 VecNormalize as used at train_td3.py:100-101;
 * SB3 is third-party and not vendored: algorithm per SURVEY Appendix D, parity unpinned) ----------
 * Running statistics are float64 vectors owned by the caller:
 *   obs_rms = mean[obs_dim], var[obs_dim], count      ret_rms = mean, var, count
 * (initialise mean 0, var 1, count 1e-4 like RunningMeanStd).  `scratch` is 2*obs_dim+2 doubles;
 * obs_dim is at most 56 (7 numbers for each of up to 8 facilities).
 * sn_vecnorm_update: RunningMeanStd.update(obs) into obs_rms_out; if `reward` is given also
 *   returns <- returns*gamma + reward and RunningMeanStd.update(returns) into ret_rms_out
 *   (the in and out buffers may be the same: every statistic is read before any is written).
 * sn_vecnorm_apply: obs_out <- clip((obs-mean)/sqrt(var+eps), +-clip_obs);
 *   reward_out <- clip(reward/sqrt(ret_var+eps), +-clip_reward); returns <- 0 if reset_returns.
 *   Either the obs or the reward part may be skipped by passing NULL. */
int32_t sn_vecnorm_update(int64_t n_envs, int32_t obs_dim, const float* obs, const float* reward, double gamma,
                          double* returns, const double* obs_rms_in, double* obs_rms_out,
                          const double* ret_rms_in, double* ret_rms_out, double* scratch, void* stream);
int32_t sn_vecnorm_apply(int64_t n_envs, int32_t obs_dim, const float* obs, float* obs_out, const float* reward,
                         float* reward_out, const double* obs_rms, const double* ret_rms, float clip_obs,
                         float clip_reward, double epsilon, int32_t reset_returns, double* returns, void* stream);

/* ---- device-side table generation for throughput runs -------------------------------------------------
 * Fill `count` elements of a demand / init table (dtype as everywhere) with the reference's distributions,
 * drawn from a Philox-4x32-10 counter stream (seed, offset + element/4): uniform integers in
 * [low, high_exclusive) like np.random.randint (utils/demands.py:44; network_components.py:143,157,297), or
 * round(max(mean + sd*N(0,1), 0)) (utils/demands.py:47).  Same distributions as the reference, NOT numpy's
 * MT19937 streams: parity runs feed host-generated tables.  dst must be 16-byte aligned. */
int32_t sn_fill_uniform(int32_t dtype, void* dst, int64_t count, int32_t low, int32_t high_exclusive, uint64_t seed,
                        uint64_t offset, void* stream);
int32_t sn_fill_normal_demand(int32_t dtype, void* dst, int64_t count, double mean, double sd, uint64_t seed,
                              uint64_t offset, void* stream);
/* The same fills with the stream position in device memory: element i is drawn at counter offset *counter + offset + i/4
 * (counter: device uint64, may be NULL = 0).  Identical launches then draw fresh tables at every replay of a captured
 * CUDA graph; sn_counter_add advances the position (one thread) between replays. */
int32_t sn_fill_uniform_at(int32_t dtype, void* dst, int64_t count, int32_t low, int32_t high_exclusive, uint64_t seed,
                           const uint64_t* counter, uint64_t offset, void* stream);
int32_t sn_fill_normal_demand_at(int32_t dtype, void* dst, int64_t count, double mean, double sd, uint64_t seed,
                                 const uint64_t* counter, uint64_t offset, void* stream);
int32_t sn_counter_add(uint64_t* counter, uint64_t delta, void* stream);

/* ---- episode tables: row-major per-env arrays -> the structure-of-arrays rows the kernels read ----------------
 * What the reference's reset leaves per env (initial inventories / pipelines, network_components.py:138-168,
 * 293-319; the demand stream, utils/demands.py:37-67) arrives as [n_envs][width] arrays (any of the four source
 * types, src_stride elements between envs).  sn_pack_rows transposes and converts them into `rows` rows of `ld`
 * elements (dst_row_stride apart; demand tables of trees interleave the sources, hence the stride), zero-filling
 * the padding columns n_envs..ld-1 and the rows width..rows-1.  One launch per table. */
enum { SN_SRC_I32 = 0, SN_SRC_I64 = 1, SN_SRC_F32 = 2, SN_SRC_F64 = 3 };
int32_t sn_pack_rows(int32_t src_dtype, const void* src, int64_t n_envs, int32_t width, int64_t src_stride, int32_t dtype,
                     void* dst, int32_t rows, int64_t ld, int64_t dst_row_stride, void* stream);

/* ---- replay buffer of the off-policy trainers (SB3 ReplayBuffer.add / .sample behind train_td3.py:109-123) --------
 * Five caller-owned float32 arrays of `capacity` rows; the write position, the number of rows stored and the
 * sampler's stream position live in DEVICE memory, so that one captured launch serves every step of a training loop
 * whose collection and updates are replayed CUDA graphs.
 *   sn_replay_add     rows (*pos + r) % capacity <- (obs[r], next_obs[r], act[r], rew[r], done[r]), r = 0..n-1 (done NULL =
 *                     zeros); obs_carry (NULL or [n][obs_dim], may be `obs` itself) receives next_obs - the loop's "current
 *                     observation" for the next step; then *pos <- (*pos + n) % capacity, *size <- min(*size + n, capacity).
 *   sn_replay_sample  row b of the batch = stored row floor(u_b * *size), u_b uniform in [0, 1) drawn from the Philox
 *                     stream (seed, *counter + b); then *counter += batch.  Uniform with replacement, like SB3. */
typedef struct sn_replay {
  float* obs;          /* [capacity][obs_dim] */
  float* next_obs;     /* [capacity][obs_dim] */
  float* act;          /* [capacity][act_dim] */
  float* rew;          /* [capacity]          */
  float* done;         /* [capacity]          */
  int64_t capacity;
  int32_t obs_dim, act_dim;
  int64_t* pos;        /* device: next row to write       */
  double* size;        /* device: rows stored so far (<= capacity) */
} sn_replay;
int32_t sn_replay_add(const sn_replay* rb, int64_t n, const float* obs, const float* next_obs, const float* act,
                      const float* rew, const float* done, float* obs_carry, void* stream);
int32_t sn_replay_sample(const sn_replay* rb, int64_t batch, uint64_t seed, uint64_t* counter, float* obs, float* next_obs,
                         float* act, float* rew, float* done, void* stream);

/* ---- the path's only collective: all-reduce (sum) of the statistics vector, ONE kernel over peer memory ------------
 * Ranks (one process per GPU, envs sharded by index) exchange nothing but the SN_STATS_WORDS doubles that sn_step /
 * sn_rollout / sn_episode accumulate (the reference has no multi-device path: train_td3.py:82,100 builds
 * n_env = 2 sequential envs in one process).  Every rank owns a MAILBOX in device memory that all peers map over
 * NVLink; sn_stats_allreduce stores this rank's vector into every rank's mailbox (16-byte entries carrying value and
 * sequence tag together), waits until its own mailbox holds all ranks' vectors of this call, and writes their sum in
 * rank order (bit-identical on every rank) to stats_out (may be stats_in).  One launch of 16 x world threads on the
 * caller's stream; no library stream, no copy, nothing to synchronise on the host.
 *   sn_mail_bytes(world)          size of one mailbox (2 parities x world slots x 16 entries of 16 bytes)
 *   sn_mail_alloc(world, &mail, handle)   cudaMalloc + zero fill on the CURRENT device; handle: SN_MAIL_HANDLE_BYTES
 *                                 bytes to send to the peers (all zero if the memory cannot be exported: then it
 *                                 serves ranks inside this process only, see sn_enable_peer_access)
 *   sn_mail_open(handle, &peer)   maps a peer's mailbox into this process (peer access enabled by the driver)
 *   sn_mail_close / sn_mail_free  release what sn_mail_open / sn_mail_alloc returned
 *   sn_enable_peer_access(d, p)   one process driving several GPUs: lets kernels on device d write to device p
 *   sn_stats_allreduce(in, out, mails, rank, world, seq, stream)
 *                                 mails: DEVICE array of `world` mailbox pointers as mapped in this process
 *                                 (entry `rank` = the own mailbox); seq = 1, 2, 3, ... the same on every rank, one
 *                                 per call; all ranks make the same calls in the same order, each on one stream.
 *                                 A peer that does not arrive within about half a minute traps the kernel. */
#define SN_MAIL_MAX_RANKS 16
#define SN_MAIL_HANDLE_BYTES 64
int64_t sn_mail_bytes(int32_t world);
int32_t sn_mail_alloc(int32_t world, void** mail, unsigned char* handle);
int32_t sn_mail_open(const unsigned char* handle, void** peer_mail);
int32_t sn_mail_close(void* peer_mail);
int32_t sn_mail_free(void* mail);
int32_t sn_enable_peer_access(int32_t device, int32_t peer);
int32_t sn_stats_allreduce(const double* stats_in, double* stats_out, void* const* mails, int32_t rank, int32_t world,
                           uint32_t seq, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SUPPLYNET_H */
