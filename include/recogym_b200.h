/*
 * recogym_b200.h -- C ABI of the B200-native reco-gym-v1 simulation hot path.
 *
 * The reference (criteo-research/reco-gym) has no FFI: its hot path is a Python object protocol
 * (SURVEY.md 8b).  This header is the native boundary a maintainer of the reference would bind
 * (ctypes stub in INTEGRATION.md); every entry point names the reference interface it replaces
 * (paths relative to the reference checkout).  Plain pointers and sizes only, no torch types.
 *
 * Conventions
 *   - every function returns RG_OK (0) or a negative RG_E* code; rg_last_error() gives the text
 *     (thread-local).  Nothing aborts the process (reference: bare asserts -> AssertionError).
 *   - "host" entry points take HOST buffers and perform the host<->device copies themselves;
 *     "_device" entry points take DEVICE buffers (caller-owned, e.g. torch tensors) and launch
 *     asynchronously on `stream` (a cudaStream_t passed as void*; NULL = default stream).
 *   - ids are int32/int64 (the reference's int16/UInt16 casts, reco_env_v1.py:123 and
 *     abstract.py:214,319-321, overflow beyond P=32767 / 65535 users; widened on purpose).
 *   - all discrete results are defined by the RNG spec in DESIGN.md: Philox4x32-10,
 *     key = seed, counter = (user_lo, user_hi, event_time, slot).
 */
#ifndef RECOGYM_B200_H
#define RECOGYM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RG_OK 0
#define RG_EINVAL (-1)   /* bad argument (shape, null pointer, out-of-range probability) */
#define RG_ECUDA (-2)    /* CUDA runtime error; text in rg_last_error() */
#define RG_ENOMEM (-3)
#define RG_ENODEVICE (-4) /* no CUDA device / wrong architecture: there is NO CPU fallback */

/* Markov states, recogym/envs/abstract.py:41-43 */
#define RG_ORGANIC 0
#define RG_BANDIT 1
#define RG_STOP 2

/* device policies for the fused rollout */
#define RG_POLICY_UNIFORM 0 /* default logging policy, abstract.py:211-221 (uniform over products) */
#define RG_POLICY_TABLE 1   /* action = table[last viewed product]: OrganicCount.act (agents/organic_count.py:54-72),
                               BanditCount.act (agents/bandit_count.py:28-47), BanditMFSquare.act (agents/bandit_mf.py:44-87) */
#define RG_POLICY_USER_VIEWS 2 /* OrganicUserEventCounterAgent.act with its default configuration
                               (agents/organic_user_count.py:29-116: exploit_explore, epsilon = 0, select_randomly):
                               action ~ Categorical(views_of_this_user / their sum) by numpy's inverse CDF
                               (cumsum, /= cdf[-1], searchsorted right) on the policy stream's second uniform.
                               The per-user view counts (ViewsFeaturesProvider, agents/abstract.py:318-409) live in
                               a device scratch owned by rg_env. */
#define RG_POLICY_EPS_TABLE 3  /* EpsilonGreedy around a table agent (agents/epsilon_greedy.py:21-71, epsilon_pure_new,
                                  not epsilon_select_worse): policy stream words 0,1 = the uniform of the coin
                                  choice([True, False], p = [eps, 1 - eps]); exploit -> table[last view]; explore ->
                                  words 2,3 = the uniform of choice(P, p = ones with a 0 at the greedy action / sum).
                                  An action differs from table[last view] exactly when it was an exploring one. */

/* math modes of the organic-view sampler */
#define RG_MATH_EXACT 0 /* every view sampled in binary64 (reference arithmetic) */
#define RG_MATH_FAST 1  /* fp32 CUDA-core contraction with a rigorous ambiguity filter; ambiguous
                           events are re-sampled in binary64, so discrete results equal RG_MATH_EXACT */
#define RG_MATH_TENSOR 2 /* as RG_MATH_FAST with the contraction on the tensor cores (tcgen05, fp16 hi/lo operand split,
                            fp32 accumulators in TMEM); the filter bound for the tensor-core sums is empirical
                            (DESIGN.md).  Operands outside the fp16 range: Gamma / mu_organic -> the CUDA-core
                            contraction runs instead (same results); omega or the shift of a user at run time ->
                            counted in totals[9], which must be 0 */

/* Static parameters of one environment == what RecoEnv1.set_static_params() builds
 * (recogym/envs/reco_env_v1.py:51-75) plus the config scalars update_state()/reset() read. */
typedef struct rg_env_desc {
    int32_t num_products;            /* config.num_products */
    int32_t K;                       /* config.K */
    const double *Gamma;             /* host [P][K] row-major   (reco_env_v1.py:64-66) */
    const double *mu_organic;        /* host [P]                (reco_env_v1.py:69-72) */
    const double *beta;              /* host [P][K] row-major   (reco_env_v1.py:133-174) */
    const double *mu_bandit;         /* host [P] */
    double state_transition[9];      /* row-major 3x3           (reco_env_v1.py:54-61) */
    double sigma_omega_initial;      /* reco_env_v1.py:80-82 */
    double sigma_omega;              /* reco_env_v1.py:96-99 */
    int32_t change_omega_for_bandits;/* reco_env_v1.py:95,106 */
    int32_t reserved;
    uint64_t seed;                   /* config.random_seed + epoch (abstract.py:59-62) */
    int32_t time_generator;          /* 0 = DefaultTimeGenerator (default_time_generator.py:4-16: the clock is the event
                                        index); 1 = NormalTimeGenerator (normal_time_generator.py:7-31: the clock advances
                                        by |N(mu, sigma)| per event, the omega drift is scaled by that step,
                                        reco_env_v1.py:89-99; the step of event index i is drawn from Philox slot 2 at
                                        (user, i)) */
    int32_t reserved2;
    double normal_time_mu;           /* config.normal_time_mu (default 0) */
    double normal_time_sigma;        /* config.normal_time_sigma (default 1) */
} rg_env_desc;

typedef struct rg_env rg_env; /* opaque: owns the device copies of the static parameters */

/* AbstractEnv.init_gym + RecoEnv1.set_static_params (abstract.py:64-88): uploads the parameters
 * to `device` (f64 row-major beta, f64 K-major Gamma, fp32 copies for RG_MATH_FAST). */
int rg_env_create(const rg_env_desc *desc, int device, rg_env **out);
/* AbstractEnv.reset_random_seed(epoch) (abstract.py:59-62): seed = random_seed + epoch */
int rg_env_set_seed(rg_env *env, uint64_t seed);
int rg_env_destroy(rg_env *env);

typedef struct rg_policy {
    int32_t kind;          /* RG_POLICY_* */
    int32_t reserved;
    const int32_t *table;  /* [P] (RG_POLICY_TABLE, RG_POLICY_EPS_TABLE); host pointer for rg_rollout_host,
                              device pointer for rg_rollout_device */
    double epsilon;        /* RG_POLICY_EPS_TABLE: exploration probability in (0, 1); needs P >= 2 */
} rg_policy;

/* Outputs of a rollout; every pointer is nullable.  Per-user arrays are indexed by
 * (user_id - user_begin) and have n_organic_users + n_users entries. */
typedef struct rg_rollout_out {
    int32_t *n_events;   /* update_state() calls of the user == its final current_time */
    int32_t *n_organic;  /* organic views */
    int32_t *n_bandit;   /* delivered recommendations (phantom row excluded) */
    int32_t *n_clicks;
    int32_t *last_view;
    uint64_t *hash;      /* FNV-1a chain over the user's event records (DESIGN.md "event hash") */
    double *omega;       /* [users][K] final latent state */
    int64_t *view_hist;  /* [P] accumulated */
    int64_t *action_hist;/* [P] accumulated (delivered actions) */
    int64_t *click_hist; /* [P] accumulated (clicked actions; evaluate_agent.py:121-124) */
    int64_t *totals;     /* [16] accumulated: [0] users, [1] events, [2] organic, [3] bandit, [4] clicks,
                            [5] phantom rows, [6] draws the fp32 filter sent to tier 2, [7] draws whose uniform lay within
                            binary64 rounding distance of a CDF boundary and were therefore re-drawn in the oracle's own operation
                            order (tier 4; ~1e-7 of the views; correct by construction, reported for information), [8] draws tier 2 sent to tier 3, [9] users whose
                            results are invalid (must be 0): RG_POLICY_USER_VIEWS users whose distinct views overflowed the
                            device list, RG_MATH_TENSOR users whose omega left the fp16 operand range (|x| >= 65000); [10..15] diagnostic, CTA-batched
                            kernels: SM cycles per phase (compact, pass 1, pass 2, re-sampling, advance, omega) */
    /* Event log == the rows of the generate_logs DataFrame (abstract.py:256-327), unordered: one int32x4
     * record {u = user_id - user_begin, t, x = product viewed | action, flags} per row, flags bit 0: z
     * (0 organic, 1 bandit), bit 1: click, bit 2: phantom row after done (abstract.py:311-316).  Rows of one
     * user have distinct t, so sorting by (u, t) restores the reference order.  *log_count accumulates the
     * rows produced; rows beyond log_capacity are dropped (call again with a larger buffer). */
    int32_t *log;        /* [log_capacity][4] or NULL */
    int64_t log_capacity;
    int64_t *log_count;  /* [1] */
} rg_rollout_out;

/* AbstractEnv.generate_logs(num_offline_users=n_users, agent, num_organic_offline_users=
 * n_organic_users) (abstract.py:241-327) reduced to counters: users user_begin.. are simulated to
 * the end of their episode with the policy acting on device.  The first n_organic_users ids are
 * organic-only (abstract.py:293-297).  test_agent's and verify_agents' success/failure counts
 * (bench_agents.py:203-206, evaluate_agent.py:742-745) are totals[4] and totals[3]+totals[5]-totals[4]. */
int rg_rollout_device(rg_env *env, const rg_policy *policy, int64_t user_begin, int64_t n_organic_users,
                      int64_t n_users, const rg_rollout_out *out_device, int math_mode, void *stream);
/* Same through HOST buffers: table H2D, outputs D2H, synchronous. */
int rg_rollout_host(rg_env *env, const rg_policy *policy, int64_t user_begin, int64_t n_organic_users,
                    int64_t n_users, const rg_rollout_out *out_host, int math_mode);

/* ---- Ordered, columnar logs and the reductions over them -------------------------
 * All pointers are DEVICE memory.  A column that is NULL is skipped. */
typedef struct rg_log_columns {
    int64_t capacity;  /* rows each column can hold */
    float *t;          /* 't'  float32 (abstract.py:318) */
    int32_t *u;        /* 'u'  user id (int32; the reference's UInt16 overflows at 65 535 users, abstract.py:319) */
    int8_t *z;         /* 'z'  0 = 'organic', 1 = 'bandit' */
    int32_t *v;        /* 'v'  viewed product, -1 where the reference has NaN */
    int32_t *a;        /* 'a'  action, -1 where the reference has NaN */
    float *c;          /* 'c'  reward (0/1; the phantom row after done carries 0), NaN for organic rows */
    double *ps;        /* 'ps' propensity of the logged action (NaN for organic rows and for RG_POLICY_USER_VIEWS) */
    int32_t *lv;       /* derived: the acting last view, i.e. Agent.update_lpv's value when the row's action was chosen */
    int32_t *lvp;      /* derived: lv of the PREVIOUS train call == BanditCount.train's row index
                          (agents/bandit_count.py:54-57); -1 = None (numpy then updates the whole row `a`); organic rows -1 */
    int32_t *sess;     /* derived: rows since the first view of the organic session this view belongs to */
    int8_t *flags;     /* derived: 1 = phantom row after done (abstract.py:311-316), 2 = view of the user's final
                          session (after its last real bandit row), 4 = the user has no real bandit row */
} rg_log_columns;

/* Which observations a training loop of the reference hands to Agent.train: */
#define RG_TRAIN_TEST_AGENT 0   /* test_agent (bench_agents.py:169-190): every session, every bandit row incl. phantom */
#define RG_TRAIN_EVOLUTION_STEP 1 /* evaluate_agent's step loop (evaluate_agent.py:100-139): stops at done -- no phantom
                                     call, the final session is never trained on */
#define RG_TRAIN_EVOLUTION_INIT 2 /* evaluate_agent's warm-up (evaluate_agent.py:56-78): as STEP, but a user whose first
                                     observation is already done gets the phantom call */

/* AbstractEnv.generate_logs (abstract.py:241-327), complete and in the reference's row order, on the device:
 * fused rollout -> exclusive scan of the per-user row counts -> one scatter (position = row_offset[u] + t, no sort) ->
 * column writer.  row_offset_device [n_organic_users + n_users + 1] and last_view_device [n_organic_users + n_users]
 * are optional outputs; rows_device[0] receives the number of rows of the log -- if it exceeds cols->capacity the tail
 * was dropped: call again with larger columns (the rows are a pure function of the arguments).
 * ps_table_device (nullable, [P]): 'ps' reported by a table agent for each last view (BanditCount, BanditMFSquare).
 * prev_last_view: BanditCount.last_product_viewed before the first user of this call (-1 = None); train_mode
 * (RG_TRAIN_*) decides which train calls move it, i.e. what the `lvp` column holds; carry_device (nullable, int32[1])
 * receives its value after the last user of this call.
 * totals_device as rg_rollout_out.totals.  Temporary buffers come from the stream-ordered allocator. */
int rg_generate_logs_device(rg_env *env, const rg_policy *policy, const double *ps_table_device, int64_t user_begin,
                            int64_t n_organic_users, int64_t n_users, int32_t prev_last_view, int train_mode,
                            int32_t *carry_device, const rg_log_columns *cols_device, int64_t *row_offset_device, int32_t *last_view_device,
                            int64_t *rows_device, int64_t *totals_device, int math_mode, void *stream);

/* Agent.train over a whole log, as scatter-adds of the count agents' sufficient statistics (all int64 [P][P], nullable,
 * accumulated; the reference holds them as float64): co_counts += b b^T per organic session (OrganicCount.train,
 * agents/organic_count.py:74-82); pulls[lvp][a] += 1, clicks[lvp][a] += c per bandit row (BanditCount.train,
 * agents/bandit_count.py:49-62).  train_mode must be the one the columns were generated with.
 * Needs the columns z, v, a, c, lvp, sess, flags. */
int rg_count_stats_device(const rg_log_columns *cols_device, int64_t rows, int32_t P, int train_mode,
                          int64_t *co_counts, int64_t *pulls, int64_t *clicks, void *stream);
/* The acting table of a count agent: kind 0 = OrganicCount.act, table[v] = argmax_j A[v][j] (A = co_counts, B unused);
 * kind 1 = BanditCount.act, table[v] = argmax_j (B[v][j] + 1) / (A[v][j] + 2) (A = pulls, B = clicks), ps[v] = that
 * ratio.  First maximum wins (numpy.argmax). */
int rg_count_policy_device(int kind, int32_t P, const int64_t *A, const int64_t *B, int32_t *table, double *ps,
                           void *stream);

/* The same statistics for num_products too large for [P][P] (BASELINE configs[3]: P = 100 000 -> 3 x 80 GB dense): a
 * device hash table, open addressing with linear probing, key = row * 2^32 + col.  All pointers are DEVICE memory owned by
 * the caller: keys[capacity] initialised to 0xFF bytes (empty), a / b [capacity] and row_a / row_b [P] and stats[2] to 0. */
typedef struct rg_sparse_counts {
    int64_t capacity;   /* slots, a power of two */
    uint64_t *keys;
    int64_t *a;         /* co_counts | pulls */
    int64_t *b;         /* (unused)  | clicks */
    int64_t *row_a;     /* BanditCount only, [P]: pulls added to EVERY cell of a row -- numpy's `pulls_a[None, a] += 1` while */
    int64_t *row_b;     /*   last_product_viewed is still None (agents/bandit_count.py:54-57) -- and the clicks likewise */
    int64_t *stats;     /* [0] occupied slots, [1] failures: insertions that found the table full / rows the policy could
                           not resolve (must stay 0; grow the table and train again otherwise) */
} rg_sparse_counts;
/* rg_count_stats_device into hash tables (either may be NULL). */
int rg_sparse_count_stats_device(const rg_log_columns *cols_device, int64_t rows, int train_mode,
                                 const rg_sparse_counts *co_counts, const rg_sparse_counts *bandit, void *stream);
/* Growth: adds every occupied cell of `from` to `to` (an empty, larger table); the caller copies row_a / row_b. */
int rg_sparse_rehash_device(const rg_sparse_counts *from, const rg_sparse_counts *to, void *stream);
/* rg_count_policy_device from a hash table: identical tables (numpy.argmax over the dense rows, first maximum wins;
 * untouched cells count 0 for kind 0 and (row_b + 1) / (row_a + 2) for kind 1). */
int rg_sparse_count_policy_device(int kind, int32_t P, const rg_sparse_counts *counts, int32_t *table, double *ps,
                                  void *stream);
/* evaluate_IPS / evaluate_SNIPS (evaluate_agent.py:753-835) of a deterministic table policy over the bandit rows of the
 * users < u_end: out4_device = {sum ee, sum ee^2, sum pi/ps, rows} with ee = c * [a == target[lv]] / ps.
 * Needs the columns z, u, a, c, ps, lv. */
int rg_ips_device(const rg_log_columns *cols_device, int64_t rows, int64_t u_end, const int32_t *target_table,
                  double *out4_device, void *stream);
/* evaluate_recall_at_k (evaluate_agent.py:852-908) of a deterministic table policy: over the bandit rows of the users
 * < u_end that were not clicked and are directly followed by an organic row of the same user, out2_device = {hits, rows}
 * with hit = [next view in topk_table[target[lv]][0..k)].  topk_table[g] = the k indices numpy.argpartition(one-hot(g), -k)
 * [-k:] returns (the reference's "top k" of a one-hot 'ps-a'), tabulated by the caller.  Needs the columns z, u, v, c, lv. */
int rg_recall_device(const rg_log_columns *cols_device, int64_t rows, int64_t u_end, const int32_t *target_table,
                     const int32_t *topk_table, int32_t k, int64_t *out2_device, void *stream);

/* Batched gym protocol.  State of N concurrent users is a caller-owned SoA on the device:
 *   user_id[N] int64, state[N] int32 (RG_*), t[N] int32 (current_time), last_view[N] int32,
 *   omega[N][K] double.
 * Organic sessions are returned as sess_v[N][sess_cap] + sess_len[N]; a user whose session did not
 * fit keeps state RG_ORGANIC and is counted in *n_unfinished (device int32, accumulated): call
 * rg_batch_continue_device to obtain the rest (appended semantics: the caller concatenates). */

/* env.reset(user_id) + env.step(None) (abstract.py:90-103,157-172; reco_env_v1.py:78-82) */
int rg_batch_reset_device(rg_env *env, int64_t N, const int64_t *user_id, int32_t *state, int32_t *t,
                          int32_t *last_view, double *omega, int32_t *sess_v, int32_t *sess_len,
                          int32_t sess_cap, int32_t *n_unfinished, int math_mode, void *stream);
/* env.step(action) (abstract.py:174-197): reward[N] = click; users with state RG_STOP are left
 * untouched (reward 0, empty session). done == (state == RG_STOP). */
int rg_batch_step_device(rg_env *env, int64_t N, const int64_t *user_id, const int32_t *action, int32_t *state,
                         int32_t *t, int32_t *last_view, double *omega, int32_t *reward, int32_t *sess_v,
                         int32_t *sess_len, int32_t sess_cap, int32_t *n_unfinished, int math_mode,
                         void *stream);
/* continues generate_organic_sessions (abstract.py:105-121) for users still RG_ORGANIC */
int rg_batch_continue_device(rg_env *env, int64_t N, const int64_t *user_id, int32_t *state, int32_t *t,
                             int32_t *last_view, double *omega, int32_t *sess_v, int32_t *sess_len,
                             int32_t sess_cap, int32_t *n_unfinished, int math_mode, void *stream);

/* Scalar spec functions evaluated on the HOST by the same source the kernels compile
 * (no GPU needed): used by the CPU tests to pin the RNG/math spec. */
void rg_spec_philox(uint64_t seed, uint64_t user, uint32_t t, uint32_t slot, uint32_t out[4]);
double rg_spec_exp(double x);
double rg_spec_log(double x);
double rg_spec_norm_ppf(double p);
double rg_spec_ff(double x);            /* reco_env_v1.py:38-41 */
double rg_spec_event_uniform(uint64_t seed, uint64_t user, uint32_t t, int which);
void rg_spec_normals(uint64_t seed, uint64_t user, uint32_t t, int K, double *z);
int32_t rg_spec_policy_uniform_action(uint64_t seed, uint64_t user, uint32_t t, int32_t P);
/* NormalTimeGenerator.new_time() (normal_time_generator.py:25-28): |mu + sigma z| of this user's `call`-th call */
double rg_spec_time_step(uint64_t seed, uint64_t user, uint32_t call, double mu, double sigma);

/* The same functions evaluated ON THE DEVICE (device buffers): kind 0 exp, 1 log, 2 norm_ppf, 3 ff map
 * in[n] -> out[n]; kinds 4 (first normal pair), 5 (event uniforms), 6 (uniform action for P=1000003, word 3)
 * map in[3n] = (seed, user, t) triples, each exactly representable in a double, to out[2n]. */
int rg_spec_eval_device(int kind, int64_t n, const double *in_device, double *out_device, void *stream);

/* Test hook of the tensor-core contraction: shifted log2-logits of 128 users against every product,
 * out[128][ceil(P/128)*128] = log2(e)*(Gamma omega + mu_organic) - shift (tcgen05, fp16 hi/lo split); the padding
 * products beyond P come out below -5e4.
 * omega: device [128][K] double, shift: device [128] float (log2 units). */
int rg_tc_logits_device(rg_env *env, const double *omega_device, const float *shift_device, float *out_device,
                        void *stream);

/* Test hook: on != 0 sends every draw that reaches a binary64 sampler (all draws of RG_MATH_EXACT, tier 3 of the fast
 * modes) through the oracle-order re-draw (tier 4), so that small runs exercise the path that otherwise serves ~1e-7 of
 * the draws.  Process-wide; returns the previous value.  (oracle/rg_oracle.c user_product_view = reco_env_v1.py:119-128.) */
int rg_debug_force_tier4(int on);

/* ---------------------------------------------------------------------------------------------------------------
 * reco-gym-v0, the toy environment (recogym/envs/reco_env_v0.py:17-69 on the AbstractEnv loops of abstract.py:90-197):
 * no latent state -- the next organic view is uniform inside the current view's product cluster, a click is
 * Bernoulli(click_probs[action, current view]); the clock never advances (every row of its logs has t = 0).
 *
 * RNG spec (same counter RNG as v1, keyed by the EVENT INDEX e = update_state() calls of the user so far):
 *   slot 3, counter 0, words 0,1: the uniform of reset()'s first product view, choice(P, p=initial_product_probs) (:55);
 *   slot 0 (event), counter e: words 0,1: update_product_view's choice(P, p=product_transition[view]) (:68-69) of an
 *     organic event / draw_click's binomial(1, p) (:64-65) of a bandit event; words 2,3: update_state's choice(3, p) (:60);
 *   slot 1 (policy), counter e: word 0 -> the uniform logging policy's action (abstract.py:214), as in v1.
 * The host builds the tables with NumPy exactly as the reference does (so that they equal it bit for bit) and the
 * device only searches them: cumulative distributions as numpy.random.RandomState.choice normalises them (cumsum,
 * divide by the last entry, searchsorted side='right'), click thresholds as numpy's legacy binomial inversion computes
 * them for n = 1 (click iff U > exp(log(1-p)) for p <= 0.5; the mirrored form for p > 0.5, stored negated). */
typedef struct rg_env0 rg_env0;
typedef struct rg_env0_desc {
    int32_t num_products;       /* P */
    int32_t num_clusters;       /* C: rows of view_cdf */
    const double *init_cdf;     /* host [P]: CDF of initial_product_probs (reco_env_v0.py:46-47) */
    const double *view_cdf;     /* host [C][P]: CDF of product_transition[v] for the products v of each cluster (:31-36) */
    const int32_t *cluster_of;  /* host [P] */
    const double *click_thr;    /* host [P][P], [action][view]: click iff U > thr (thr >= 0) / iff !(U > -thr) (thr < 0) */
    double trans_cdf[9];        /* [state][3]: CDF of state_transition[state] (:22-29), as choice() normalises it */
    uint64_t seed;
    int32_t device;
    int32_t reserved;
} rg_env0_desc;

#define RG0_POLICY_UNIFORM 0 /* abstract.py:214 */
#define RG0_POLICY_TABLE 1   /* a = table[current view] */
#define RG0_POLICY_REPLAY 2  /* ONE user (n_users = 1): the actions are given; the rollout stops at the first bandit event
                                without one (state_out = 1: the env waits for step(action)).  Every gym step() of the
                                facade is such a replay: a trajectory is a pure function of (seed, user, actions so far) */

typedef struct rg_env0_out {     /* all nullable except totals; *_host entry points take host pointers, *_device device ones */
    int64_t *totals;             /* [8] accumulated: users, events, organic, bandit, clicks, phantom rows, 0, 0 */
    int32_t *n_events, *n_organic, *n_bandit, *n_clicks, *last_view, *state; /* per user [n]; state: 2 done, 1 waiting */
    int32_t *log;                /* [log_capacity][4]: {user - user_begin, event index, product viewed | action,
                                    flags: bit 0 bandit, bit 1 click, bit 2 phantom row after done (abstract.py:311-316)} */
    int64_t log_capacity;
    int64_t *log_count;          /* rows produced (rows beyond the capacity are dropped: call again) */
} rg_env0_out;

int rg_env0_create(const rg_env0_desc *desc, rg_env0 **out);
void rg_env0_destroy(rg_env0 *env);
int rg_env0_set_seed(rg_env0 *env, uint64_t seed);   /* reset_random_seed(epoch) (abstract.py:59-62) */
/* generate_logs(n_users, agent, n_organic_users) of users user_begin.. (abstract.py:241-327) in one launch: organic-only
 * users first (their first session only), then regular users with the phantom bandit row after done. */
int rg_env0_rollout_device(rg_env0 *env, int policy_kind, const int32_t *table_or_actions_device, int64_t n_table,
                           int64_t user_begin, int64_t n_organic_users, int64_t n_users, const rg_env0_out *out_device,
                           void *stream);
int rg_env0_rollout_host(rg_env0 *env, int policy_kind, const int32_t *table_or_actions_host, int64_t n_table,
                         int64_t user_begin, int64_t n_organic_users, int64_t n_users, const rg_env0_out *out_host);
/* Test hook (CPU tests only; nothing in the package calls it): the per-user body of the kernel instantiated for the host
 * and run over the users in a loop, on the descriptor's host tables. */
int rg_spec_env0_rollout(const rg_env0_desc *desc, int policy_kind, const int32_t *table_or_actions_host, int64_t n_table,
                         int64_t user_begin, int64_t n_organic_users, int64_t n_users, const rg_env0_out *out_host);

int rg_device_count(void);              /* 0 when no usable GPU: callers must fail, not fall back */
int rg_version(void);
const char *rg_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* RECOGYM_B200_H */
