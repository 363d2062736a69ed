/*
 * beergame_c.c - plain-C restatement of the reference's serial beer-game step.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT.  This is the second form of the CPU oracle (the first is
 * oracle/beergame_np.py).  It exists so that (a) the oracle can check large batches in seconds
 * and (b) bench.py has a compiled CPU baseline ("port") to time on the GPU box's host cores.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load it.  The product (multi_echelon_drl_b200) never links or calls it.
 *
 * Parity status: PINNED - tests/test_oracle_c.py checks it against the golden trajectories
 * recorded from the unmodified reference (tests/golden/rollouts.npz, classic.npz) and against
 * beergame_np.py on random batches.
 *
 * Reference lines restated (all under /root/reference):
 *   reset   network_components.py:138-168 (Arc.reset), :293-319 (Node.reset), envs.py:64-80
 *   before  supplynetwork.py:182-210          place  network_components.py:353-383
 *   sweep   supplynetwork.py:229-268, network_components.py:189-234, :323-345
 *   cost    supplynetwork.py:285-310          state  supplynetwork.py:114-171
 *   policy  utils/heuristics.py:35-75         d+a    utils/wrappers.py:23-27
 *
 * One env is a small struct of 64-bit integers; the batch functions loop env-major (each env
 * runs all its periods before the next one starts) and split envs over pthreads.
 */
#include <stdint.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

#define NF 4

typedef struct bg_spec {
  int32_t n_agents;
  int32_t agents[NF];        /* node per action column (caller order) */
  int32_t global_observable;
  int32_t max_episode_steps;
  int32_t rule;              /* 0 = 'a', 1 = 'd+a' */
  double holding_cost[NF], backorder_cost[NF];
  int64_t coplayer_target[NF];
  int64_t coplayer_lb, coplayer_ub;
  int64_t dpa_offset, dpa_lb, dpa_ub;
} bg_spec;

typedef struct bg_policy {   /* BaseStockPolicy, one target per action column */
  int64_t target[NF];
  int64_t lb, ub;
  int32_t rule;
} bg_policy;

typedef struct bg_env {
  int64_t inv[NF];
  int64_t info[NF][2];       /* info[k][j]: order slip reaching the supplier after j+1 befores */
  int64_t B[NF];             /* arrived, unshipped orders per arc */
  int64_t ship[NF][2];       /* ship[k][j]: shipment arriving after j+1 sweeps */
  int64_t U[NF][5];          /* unreceived pipeline, newest first; 4 live slots at observe time */
  int64_t latest[NF + 1];
  int64_t unf_ind, cur;
  int64_t unf_post[NF];
  int32_t u_len;             /* 4 (observe time) or 5 (after ordering / terminal) */
  int32_t t, terminal;
} bg_env;

static const int INFO_LT[NF] = {2, 2, 2, 1};

static inline int64_t i64min(int64_t a, int64_t b) { return a < b ? a : b; }
static inline int64_t i64max(int64_t a, int64_t b) { return a > b ? a : b; }
static inline int64_t clip(int64_t x, int64_t lo, int64_t hi) { return i64min(i64max(x, lo), hi); }

static void before(bg_env* e) {                       /* supplynetwork.py:185-199 */
  for (int k = 0; k < NF; ++k) {
    int64_t a = e->info[k][0];
    e->info[k][0] = e->info[k][1];
    e->info[k][1] = 0;
    e->latest[k + 1] = a;
    e->B[k] += a;
    e->U[k][3] += e->U[k][4];
    e->U[k][4] = 0;
  }
  e->u_len = 4;
}

static inline int64_t unfilled(const bg_env* e, int k) { return k == 0 ? e->cur + e->unf_ind : e->B[k - 1]; }
static inline int64_t latest(const bg_env* e, int k) { return k == 0 ? e->cur : e->latest[k]; }

static int64_t base_stock(const bg_env* e, int k, int64_t target, int64_t lb, int64_t ub, int rule) {
  int64_t on_order = e->U[k][0] + e->U[k][1] + e->U[k][2] + e->U[k][3];
  int64_t q = target - (e->inv[k] + on_order - unfilled(e, k));
  if (rule == 1) q -= latest(e, k);
  return clip(q, lb, ub);
}

static void place(bg_env* e, int k, int64_t q) {      /* network_components.py:353-383 */
  if (q < 0) q = 0;
  e->info[k][INFO_LT[k] - 1] = q;
  for (int j = 4; j > 0; --j) e->U[k][j] = e->U[k][j - 1];
  e->U[k][0] = q;
}

static void reset_env(bg_env* e, const int32_t* init, const int32_t* demand) {
  static const int so_off[NF] = {4, 6, 8, 10};
  memset(e, 0, sizeof(*e));
  for (int k = 0; k < NF; ++k) {
    e->inv[k] = init[k];
    int64_t seq[8] = {0, 0, 0, 0, 0, 0, 0, 0};       /* [0]*4 + shipments + sales_orders */
    int len = 4;
    for (int j = 0; j < 2; ++j) { e->ship[k][j] = init[11 + 2 * k + j]; seq[len++] = e->ship[k][j]; }
    for (int j = 0; j < INFO_LT[k]; ++j) { e->info[k][j] = init[so_off[k] + j]; seq[len++] = e->info[k][j]; }
    for (int j = 0; j < 5; ++j) e->U[k][j] = seq[len - 1 - j];       /* [::-1][:5] */
  }
  e->cur = demand[0];
  before(e);
}

/* 7 numbers of node k (supplynetwork.py:151-169); a co-player below the first agent is seen
 * after it ordered inside before_action, i.e. with its own order in front (non-terminal only). */
static void node_state(const bg_env* e, const bg_spec* s, int lo, int k, float* o) {
  int64_t u[4] = {e->U[k][0], e->U[k][1], e->U[k][2], e->U[k][3]};
  if (k < lo && !e->terminal) {
    int64_t q = i64max(base_stock(e, k, s->coplayer_target[k], s->coplayer_lb, s->coplayer_ub, 0), 0);
    u[3] = u[2]; u[2] = u[1]; u[1] = u[0]; u[0] = q;
  }
  o[0] = (float)e->inv[k]; o[1] = (float)unfilled(e, k); o[2] = (float)latest(e, k);
  for (int j = 0; j < 4; ++j) o[3 + j] = (float)u[j];
}

static int first_agent(const bg_spec* s) {
  int lo = NF;
  for (int i = 0; i < s->n_agents; ++i) lo = s->agents[i] < lo ? s->agents[i] : lo;
  return lo;
}

static int observe(const bg_env* e, const bg_spec* s, float* o) {
  const int lo = first_agent(s);
  const int n_obs = s->global_observable ? NF : s->n_agents;
  for (int i = 0; i < n_obs; ++i) node_state(e, s, lo, s->global_observable ? i : s->agents[i], o + 7 * i);
  return 7 * n_obs;
}

/* envs.py:82-113; a[i] = action column i.  Returns the reward; cost[0..3] holding, [4..7] backorder. */
static double step_env(bg_env* e, const bg_spec* s, const int64_t* a, const int32_t* demand, double* cost) {
  int64_t q[NF];
  int slot[NF] = {-1, -1, -1, -1};
  for (int i = 0; i < s->n_agents; ++i) slot[s->agents[i]] = i;
  for (int k = 0; k < NF; ++k) {
    if (slot[k] >= 0) {
      q[k] = a[slot[k]];
      if (s->rule == 1) q[k] = clip(latest(e, k) + q[k] + s->dpa_offset, s->dpa_lb, s->dpa_ub);
    } else {
      q[k] = base_stock(e, k, s->coplayer_target[k], s->coplayer_lb, s->coplayer_ub, 0);
    }
  }
  for (int k = 0; k < NF; ++k) place(e, k, q[k]);
  e->u_len = 5;
  for (int sup = NF; sup >= 1; --sup) {                /* supplynetwork.py:241-264 */
    const int k = sup - 1;
    int64_t arr = e->ship[k][0];
    e->ship[k][0] = e->ship[k][1];
    e->ship[k][1] = 0;
    e->inv[k] += arr;
    for (int i = 4; i >= 0 && arr > 0; --i) {
      int64_t f = i64min(e->U[k][i], arr);
      e->U[k][i] -= f;
      arr -= f;
    }
    int64_t out;
    if (sup == NF) {
      out = e->B[k];
    } else {
      out = i64min(e->inv[sup], e->B[k]);
      e->inv[sup] -= out;
    }
    e->ship[k][1] += out;
    e->B[k] -= out;
    if (sup != NF) e->unf_post[sup] = e->B[k];
  }
  int64_t tot = e->cur + e->unf_ind;
  int64_t sold = i64max(0, i64min(e->inv[0], tot));
  e->inv[0] -= sold;
  e->unf_ind = tot - sold;
  e->unf_post[0] = e->unf_ind;
  e->cur = demand[e->t + 1];
  double ch = 0.0, cb = 0.0;                           /* supplynetwork.py:285-310 */
  for (int k = 0; k < NF; ++k) {
    double h = -(double)e->inv[k] * s->holding_cost[k];
    double b = -(double)e->unf_post[k] * s->backorder_cost[k];
    if (cost) { cost[k] = h; cost[NF + k] = b; }
    ch += h;
    cb += b;
  }
  e->t += 1;
  if (e->t >= s->max_episode_steps) e->terminal = 1; else before(e);
  return ch + cb;
}

/* ---------------------------------------------------------------------------------------------
 * Batch entry points (C ABI for ctypes).  init [n][19], demand [n][T+3] int32 row-major.
 * actions: int32 [n_steps][n][n_agents] or NULL (then `policy` drives the agents).
 * Outputs may be NULL: obs float [n_steps][n][obs_dim], reward double [n_steps][n],
 * ep_return double [n], obs0 float [n][obs_dim].  Returns env-steps executed. */
typedef struct bg_job {
  const bg_spec* spec;
  int64_t n, lo, hi;
  int32_t n_steps;
  const int32_t *init, *demand, *actions;
  const bg_policy* policy;
  float *obs0, *obs;
  double *reward, *ep_return;
} bg_job;

static void* run_range(void* arg) {
  const bg_job* j = (const bg_job*)arg;
  const bg_spec* spec = j->spec;
  const int T3 = spec->max_episode_steps + 3;
  const int na = spec->n_agents;
  const int od = 7 * (spec->global_observable ? NF : na);
  const int64_t n = j->n;
  for (int64_t i = j->lo; i < j->hi; ++i) {
    bg_env e;
    reset_env(&e, j->init + i * 19, j->demand + i * T3);
    if (j->obs0) observe(&e, spec, j->obs0 + i * od);
    double ret = 0.0;
    for (int t = 0; t < j->n_steps; ++t) {
      int64_t a[NF] = {0, 0, 0, 0};
      if (j->actions) {
        for (int c = 0; c < na; ++c) a[c] = j->actions[((int64_t)t * n + i) * na + c];
      } else {
        for (int c = 0; c < na; ++c)
          a[c] = base_stock(&e, spec->agents[c], j->policy->target[c], j->policy->lb, j->policy->ub, j->policy->rule);
      }
      double r = step_env(&e, spec, a, j->demand + i * T3, 0);
      ret += r;
      if (j->reward) j->reward[(int64_t)t * n + i] = r;
      if (j->obs) observe(&e, spec, j->obs + ((int64_t)t * n + i) * od);
    }
    if (j->ep_return) j->ep_return[i] = ret;
  }
  return 0;
}

int32_t bg_num_threads(void) {
  long c = sysconf(_SC_NPROCESSORS_ONLN);
  return c > 0 ? (int32_t)c : 1;
}

#define BG_MAX_THREADS 256
int64_t bg_rollout(const bg_spec* spec, int64_t n, int32_t n_steps, const int32_t* init, const int32_t* demand,
                   const int32_t* actions, const bg_policy* policy, int32_t n_threads, float* obs0, float* obs,
                   double* reward, double* ep_return) {
  if (n_threads <= 0) n_threads = bg_num_threads();
  if (n_threads > BG_MAX_THREADS) n_threads = BG_MAX_THREADS;
  if ((int64_t)n_threads > n) n_threads = (int32_t)n;
  bg_job jobs[BG_MAX_THREADS];
  pthread_t th[BG_MAX_THREADS];
  const int64_t chunk = (n + n_threads - 1) / n_threads;
  for (int t = 0; t < n_threads; ++t) {
    bg_job j = {spec, n, t * chunk, (t + 1) * chunk < n ? (t + 1) * chunk : n, n_steps, init, demand, actions,
                policy, obs0, obs, reward, ep_return};
    jobs[t] = j;
  }
  if (n_threads == 1) {
    run_range(&jobs[0]);
  } else {
    for (int t = 0; t < n_threads; ++t) pthread_create(&th[t], 0, run_range, &jobs[t]);
    for (int t = 0; t < n_threads; ++t) pthread_join(th[t], 0);
  }
  return n * (int64_t)n_steps;
}
