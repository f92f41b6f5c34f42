/*
 * at2.h - C ABI of libat2_b200: the B200 (sm_100a) implementation of autotune-v2's Monte-Carlo optimiser-evaluation
 * hot path (Hyperband / TPE / hybrids on Gamma-process loss simulations and on the closest-known-loss-function
 * approximation, collected into an EPDF-OFE).
 *
 * The reference (cristianmatache/autotune-v2) is pure Python and has no FFI of its own: the boundary this library
 * sits behind is the reference's Python class API (autotune/benchmarks, autotune/optimisers).  Each entry point below
 * names the reference code (file:line, relative to the reference root) whose work it replaces; INTEGRATION.md shows
 * the ctypes binding a maintainer of the reference would add.
 *
 * Conventions
 *   - plain C: pointers + sizes, no C++/torch types.  `stream` is a `cudaStream_t` passed as `void*` (NULL = legacy
 *     default stream).
 *   - every function returns 0 on success, a negative AT2_E* code on failure; `at2_last_error()` returns a
 *     thread-local message for the last failure on the calling thread.  Nothing throws or aborts across the ABI.
 *   - pointers named `d_*` are DEVICE pointers owned by the caller; everything else is host memory.  The library
 *     allocates nothing and keeps no global state; work is enqueued on `stream` and is asynchronous w.r.t. the host.
 *   - randomness is counter-based Philox4x32-10 keyed by (seed; global repetition id, arm id): results do not depend
 *     on grid shape, on how repetitions are sharded over GPUs, or on which entry point variant produced them.
 */
#ifndef AT2_H
#define AT2_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AT2_OK 0
#define AT2_EINVAL (-1)   /* bad argument (message says which) */
#define AT2_ECUDA (-2)    /* CUDA runtime error (message carries cudaGetErrorString) */
#define AT2_ELIMIT (-3)   /* a compile-time capacity (AT2_MAX_*) or the shared-memory budget is exceeded */

#define AT2_MAX_BRACKETS 16
#define AT2_MAX_ROUNDS 64
#define AT2_MAX_CKPTS 16
#define AT2_MAX_FAMILIES 8
#define AT2_MAX_PEERS 16     /* gather buffers of at2_hb_sim_gather_f32 */

/* test functions, reference autotune/benchmarks/opt_function_problem.py:141-147 */
#define AT2_FUNC_EGG 0
#define AT2_FUNC_BRANIN 1
#define AT2_FUNC_CAMEL 2
#define AT2_FUNC_WAVE 3
#define AT2_FUNC_RASTRIGIN 4
/* EXTENSION, no reference equivalent (the reference's Egg-holder is 2-D): the n-dimensional Egg-holder
 * sum_{i<3} [-(x_{i+1}+47) sin sqrt|x_{i+1} + x_i/2 + 47| - x_i sin sqrt|x_i - (x_{i+1}+47)|] over four coordinates,
 * each uniform on [x_min, x_max] x [y_min, y_max] alternately.  Hyperband entry points only (TPE suggestions are 2-D). */
#define AT2_FUNC_EGG4 5

#define AT2_SCHED_UNIFORM 0      /* UniformShapeFamilyScheduler,    core/shape_family_scheduler.py:68-77 */
#define AT2_SCHED_ROUND_ROBIN 1  /* RoundRobinShapeFamilyScheduler, core/shape_family_scheduler.py:50-65 */

/* Hyperband bracket/round table of one repetition.
 * Replaces the arithmetic of optimisers/sequential/hyperband_optimiser.py:66-97 (s_max, s_min, B, n, r, n_i, r_i,
 * int(n_i/eta)).  Rounds are listed bracket after bracket in execution order. */
typedef struct at2_plan {
    int32_t R;                                  /* max_iter */
    int32_t eta;
    int32_t n_brackets;
    int32_t n_rounds;                           /* total over brackets */
    int32_t n_arms;                             /* sum of bracket_n: evaluators created per repetition */
    int32_t n_ckpts;                            /* distinct int(r_i) values */
    int32_t n_surv;                             /* sum over rounds of min(keep, n_eval) */
    int32_t bracket_n[AT2_MAX_BRACKETS];        /* n of each bracket */
    int32_t bracket_first_arm[AT2_MAX_BRACKETS];
    int32_t bracket_first_round[AT2_MAX_BRACKETS + 1];
    int32_t round_n_eval[AT2_MAX_ROUNDS];       /* evaluators evaluated in the round */
    int32_t round_keep[AT2_MAX_ROUNDS];         /* int(n_i/eta) */
    int32_t round_time[AT2_MAX_ROUNDS];         /* 1-based trajectory position read: int(r_i), or R when int(r_i)
                                                   is 0 (python fs[-1]; e.g. R=729, eta=3) */
    int32_t round_ckpt[AT2_MAX_ROUNDS];         /* index of round_time in ckpt_time */
    int32_t ckpt_time[AT2_MAX_CKPTS];           /* ascending */
    double bracket_r0[AT2_MAX_BRACKETS];        /* r = R*eta**(-s) as the float the reference computes (transfer
                                                   predicates compare against it un-truncated) */
    int32_t round_res[AT2_MAX_ROUNDS];          /* int(r_i) itself, NOT wrapped (0 where round_time wrapped to R): what
                                                   the hybrids record per evaluator for their transfer predicates
                                                   (hybrid_hyperband_tpe_with_transfer.py:96-98,161-209) */
} at2_plan;

/* One shape family, core/shape_family_scheduler.py:10-17. */
typedef struct at2_family {
    double ml_agg, nec_agg, up_spik;
    double start_shift, end_shift;
    int32_t is_smooth;
    int32_t _pad;
} at2_family;

/* Simulation problem + scheduler configuration (benchmarks/opt_function_simulation_problem.py:91-118 arguments,
 * benchmarks/opt_function_problem.py:15-46 domains, experiments/run_simulation.py:31-57 constants). */
typedef struct at2_sim_config {
    int32_t func;            /* AT2_FUNC_* */
    int32_t n_families;
    int32_t scheduler;       /* AT2_SCHED_* */
    int32_t descending;      /* 0: min_or_max=min, 1: max (core/optimiser.py:52-54) */
    double x_min, x_max, y_min, y_max;
    double init_noise;
    at2_family fam[AT2_MAX_FAMILIES];
} at2_sim_config;

/* number of hyperparameters of a test function: 2 for the reference's functions, 4 for AT2_FUNC_EGG4; -1 if unknown */
int32_t at2_func_dims(int32_t func);

const char* at2_version(void);
const char* at2_last_error(void);

/* Debug options (tests and measurements only; all zero by default = the library chooses): tile / CTA shapes and code-path
 * switches that never change a result.  Process-wide; the library never reads the environment.  Names: hb_reps_per_tile,
 * hb_team_warps, hb_warps_per_cta, hb_no_split, hb_equal_tiles, hb_no_tile_halving, hb_static_schedule, hb_singles,
 * force_gtab, lazy_reps_per_tile, tpe_warps_per_cta, kf_warps_per_cta, kf_reps_per_tile; "reset" clears all. */
int at2_debug_option(const char* name, int32_t value);

/* ---- host-only helpers ------------------------------------------------------------------------------------------ */

/* Fill `plan` for Hyperband(R, eta) with the reference's float arithmetic (hyperband_optimiser.py:66-97).
 * s_max is the exact integer floor(log_eta R) (what the reference's 64-digit mpmath evaluation yields). */
int at2_bracket_plan(int32_t R, int32_t eta, at2_plan* plan);

/* Rows of the linear operator W with scipy.signal.savgol_filter(x, window, polyorder=3, mode='interp') == W @ x
 * (core/simulation_evaluator.py:76-82).  `window` = at2_savgol_window(R).  out[i*R + j] = W[rows[i]][j]. */
int32_t at2_savgol_window(int32_t R);
int at2_savgol_rows(int32_t R, int32_t n_rows, const int32_t* rows, double* out);

/* Size in bytes / contents of the per-configuration constant tables consumed by at2_hb_sim_*, at2_tpe_sim_f32 and
 * at2_sim_trajectories_f32: one record per (family, step) with the Marsaglia-Tsang constants of the step's Gamma draw
 * (core/simulation_evaluator.py:36-42) and the folded constants of the recurrence step built from the two
 * time-aggressiveness powers (t/(R-1))**nec and (t/(R-1))**(1.1*nec) (:105,:112), the Savitzky-Golay rows at the plan's
 * checkpoints, and (float) the folded-step table of the pre-drawn path.  `is_f64` selects double (parity path) or float
 * (production path) entries.  The tables BELONG TO THE PLAN they were built for: the float records flag the plan's
 * checkpoint positions (the trajectory loop stores a loss where a record is flagged) and the Savitzky-Golay columns are
 * the plan's checkpoints - pass the same plan to the build and to the launch.  The caller copies `host_out` to the device
 * and passes it as `d_tables`. */
/* Families with equal (ml_agg, nec_agg) share one block of per-step records in the tables (their other attributes are
 * per-arm scalars): slots[f] = block of family f; returns the number of blocks (<= n_families), negative on bad input. */
int32_t at2_family_slots(const at2_sim_config* cfg, int32_t* slots);
int64_t at2_hb_tables_bytes(const at2_plan* plan, const at2_sim_config* cfg, int32_t is_f64);
int at2_hb_tables_build(const at2_plan* plan, const at2_sim_config* cfg, int32_t is_f64, void* host_out);

/* ---- (a)+(b): fused simulation + successive halving -------------------------------------------------------------- */

/* Optional per-repetition outputs of at2_hb_sim_*; any pointer may be NULL.  `real` is float or double. */
typedef struct at2_hb_outputs {
    void* d_ofe;            /* real[n_reps]            min/max over round-bests (core/optimiser.py:67-68) */
    int32_t* d_ofe_arm;     /* int32[n_reps]           arm id (creation order within the repetition) of the OFE */
    void* d_best_xy;        /* real[n_reps][dims]      hyperparameters of that arm, dims = at2_func_dims(cfg->func) */
    void* d_round_best;     /* real[n_reps][n_rounds]  best-in-round (hyperband_optimiser.py:98-99) */
    int32_t* d_round_best_arm; /* int32[n_reps][n_rounds] */
    int32_t* d_survivors;   /* int32[n_reps][n_surv]   survivors of every round, sorted order, rounds concatenated */
    void* d_ckpt_loss;      /* real[n_reps][n_arms][n_ckpts] loss of every arm at every checkpoint */
} at2_hb_outputs;

/* Production path, FP32, Philox.  Replaces the repetition loop experiments/run_simulation.py:147-157 for
 * method sim(hb): per repetition draw arms (core/arm.py:44-55) and families, simulate every arm's Gamma-process
 * trajectory (core/simulation_evaluator.py:84-116, benchmarks/opt_function_simulation_problem.py:50-86), read the
 * losses at the Hyperband checkpoints and run successive halving (hyperband_optimiser.py:74-106).
 * Repetition ids rep_offset .. rep_offset+n_reps-1 key the RNG. */
int at2_hb_sim_f32(const at2_plan* plan, const at2_sim_config* cfg, const void* d_tables, int64_t n_reps,
                   int64_t rep_offset, uint64_t seed, const at2_hb_outputs* out, void* stream);

/* Targets of the fused all-gather of the OFE vector (the one collective of the path: the OFEs of all ranks feed the KDE,
 * experiments/run_simulation.py:156-157 + simulation_evaluation/plot_results.py:11-17).  The OFE of repetition i of a call is
 * additionally stored to element base + i of every target: the peers' symmetric-memory buffers mapped into this process
 * (stores travel over NVLink / NVSwitch while the kernel is still simulating), or ONE NVSwitch multicast address that fans a
 * store out to every GPU - then is_multicast must be 1: the kernel issues multimem.st (the only kind of access the PTX ISA
 * defines on a multimem address).  The caller orders the stores against the readers (a barrier across the ranks). */
typedef struct at2_gather {
    float* d_target[AT2_MAX_PEERS];
    int32_t n_targets;      /* 1..AT2_MAX_PEERS; exactly 1 when is_multicast */
    int32_t is_multicast;
    int64_t base;
} at2_gather;

/* at2_hb_sim_f32 fused with that all-gather.  Every other argument and output as in at2_hb_sim_f32; out->d_ofe may be NULL. */
int at2_hb_sim_gather_f32(const at2_plan* plan, const at2_sim_config* cfg, const void* d_tables, int64_t n_reps,
                          int64_t rep_offset, uint64_t seed, const at2_hb_outputs* out, const at2_gather* gather,
                          void* stream);

/* Same repetitions with early stopping - the product's default whenever the caller does not ask for the losses of unread
 * checkpoints: an arm is only simulated as far as the round that reads it needs (the reference simulates every arm to R on
 * its first evaluate(), benchmarks/opt_function_simulation_problem.py:67-76, but only ever reads the checkpoints - SURVEY
 * D5), and a raw arm read at R is its target f_n without any simulation (core/simulation_evaluator.py:105-107,113-115).
 * Streams and arithmetic are those of at2_hb_sim_f32, so every output is bit-identical to it; out->d_ckpt_loss must be NULL
 * (the losses of unread checkpoints are never computed).  The _gather variant adds the fused all-gather of the OFEs.
 * The launch takes a few MB to ~120 MB of stream-ordered scratch (9 bytes per arm of the tiles in flight: the arms' start
 * value, target and family, read back by later rounds) from a pool the library keeps; if the allocation is refused the
 * kernel redraws the arms every round instead - the same results, a few percent slower. */
int at2_hb_sim_early_stop_f32(const at2_plan* plan, const at2_sim_config* cfg, const void* d_tables, int64_t n_reps,
                              int64_t rep_offset, uint64_t seed, const at2_hb_outputs* out, void* stream);
int at2_hb_sim_early_stop_gather_f32(const at2_plan* plan, const at2_sim_config* cfg, const void* d_tables, int64_t n_reps,
                                     int64_t rep_offset, uint64_t seed, const at2_hb_outputs* out, const at2_gather* gather,
                                     void* stream);

/* Pre-drawn inputs (parity paths): nothing is sampled; the kernel only runs the recurrence, the checkpoint read-out
 * and successive halving.  Per arm: first and target values f_1, f_n (opt_function_simulation_problem.py:61-64),
 * the family index, and the R-1 Gamma values of its (effective) trajectory.
 *   d_f1, d_fn : real[n_reps][n_arms]      d_family : int32[n_reps][n_arms]
 *   d_gamma    : real[n_reps][n_arms][R-1]  d_xy (optional, only echoed to d_best_xy) : real[n_reps][n_arms][2]
 * The f64 variant executes the reference's float64 operation order with un-fused IEEE add/mul/div, so trajectories
 * of non-smooth families, survivor sets and OFEs are bit-identical to the reference on the same inputs. */
int at2_hb_sim_predrawn_f64(const at2_plan* plan, const at2_sim_config* cfg, const void* d_tables, int64_t n_reps,
                            const double* d_f1, const double* d_fn, const int32_t* d_family, const double* d_gamma,
                            const double* d_xy, const at2_hb_outputs* out, void* stream);
int at2_hb_sim_predrawn_f32(const at2_plan* plan, const at2_sim_config* cfg, const void* d_tables, int64_t n_reps,
                            const float* d_f1, const float* d_fn, const int32_t* d_family, const float* d_gamma,
                            const float* d_xy, const at2_hb_outputs* out, void* stream);

/* ---- (c): TPE / Hyperband-TPE hybrids on the simulation problem ----------------------------------------------------- */

#define AT2_TRANSFER_NONE 0       /* HybridHyperbandTpeOptimiser / ...NoTransferOptimiser */
#define AT2_TRANSFER_ALL 1        /* ...TransferAllOptimiser      (hybrid_hyperband_tpe_with_transfer.py:177-186) */
#define AT2_TRANSFER_LONGEST 2    /* ...TransferLongestOptimiser  (:170-174) */
#define AT2_TRANSFER_SAME 3       /* ...TransferSameOptimiser     (:189-198) */
#define AT2_TRANSFER_THRESHOLD 4  /* ...TransferThresholdOptimiser (:201-209) */

/* hyperopt.tpe.suggest parameters as the reference uses them (tpe_optimiser.py:51-53 passes only n_startup_jobs;
 * the rest are hyperopt 0.2.4 defaults: gamma 0.25, 24 candidates, prior weight 1, linear forgetting 25). */
typedef struct at2_tpe_config {
    int32_t enabled;       /* 0: first-round arms are random (plain Hyperband / random search) */
    int32_t n_startup;     /* 10; injected trials are added on top (tpe_optimiser.py:52) */
    int32_t n_candidates;  /* 24 (<= 32: one candidate per lane) */
    int32_t lf;            /* 25 */
    int32_t transfer;      /* AT2_TRANSFER_* */
    int32_t threshold;     /* n_res_transfer_threshold of the Threshold variant */
    double gamma;          /* 0.25 */
    double prior_weight;   /* 1.0 */
} at2_tpe_config;

/* The successive-halving selection alone (parity of survivor indices): for each of n_segments segments
 * d_losses[d_seg_offsets[s] .. d_seg_offsets[s+1]) write the positions (0-based within the segment) of its d_keep[s] best
 * losses, best first, to d_idx_out at the segment's own offsets; entries past keep are set to -1.  Order and tie-breaks are
 * python's stable sorted(evaluators, key=loss, reverse=descending)[:keep] (optimisers/sequential/hyperband_optimiser.py:
 * 46-57); keep is clamped to [0, segment length].  max_seg_len: an upper bound on the segment lengths (<= 65536; sizes the
 * per-warp sort scratch) - a longer segment is an error.  Runs the same warp primitives as the fused kernels (ranking by
 * counting, register / shared-memory bitonic networks).  Synchronises `stream` before returning. */
int at2_segmented_topk_f32(const float* d_losses, const int64_t* d_seg_offsets, const int32_t* d_keep, int64_t n_segments,
                           int32_t max_seg_len, int32_t descending, int32_t* d_idx_out, void* stream);
int at2_segmented_topk_f64(const double* d_losses, const int64_t* d_seg_offsets, const int32_t* d_keep, int64_t n_segments,
                           int32_t max_seg_len, int32_t descending, int32_t* d_idx_out, void* stream);

/* Plan of a stand-alone optimiser that evaluates n_evals arms once at n_resources: TpeOptimiser
 * (tpe_optimiser.py:40-60) with tpe.enabled = 1, RandomOptimiser (random_optimiser.py:39-58) with enabled = 0. */
int at2_single_round_plan(int32_t R, int32_t n_evals, int32_t n_resources, at2_plan* plan);

/* Hyperband whose first round of every bracket is a TPE run (+ history transfer), fused with simulation and halving:
 * replaces the repetition loop of experiments/run_simulation.py:147-157 for methods sim(hb+tpe), sim(hb+tpe+transfer+*),
 * sim(tpe), sim(random).  Same tables/outputs as at2_hb_sim_f32; with tpe->enabled == 0 and a Hyperband plan the results
 * are bit-identical to at2_hb_sim_f32.  d_arm_xy (optional): float[n_reps][n_arms][2] hyperparameters of every arm.
 * d_obs_loss (optional): float[n_reps][n_arms], the loss the TPE phase observed for an arm (its read-out at the bracket's
 * first resource level, tpe_optimiser.py:62-89) - written only for arms that were observed, so pre-fill with NaN.
 * d_n_injected (optional): int32[n_reps][n_brackets], how many evaluations of earlier brackets were transferred into the
 * bracket's TPE run (_get_trials, hybrid_hyperband_tpe_with_transfer.py:116-158). */
int at2_tpe_sim_f32(const at2_plan* plan, const at2_sim_config* cfg, const at2_tpe_config* tpe, const void* d_tables,
                    int64_t n_reps, int64_t rep_offset, uint64_t seed, const at2_hb_outputs* out, float* d_arm_xy,
                    float* d_obs_loss, int32_t* d_n_injected, void* stream);

/* Parzen-estimator scoring alone (parity of the estimator arithmetic): for each of n_problems independent 1-D
 * hp.uniform(low, high) problems with n_obs observations (d_obs, d_loss: float[n_problems][n_obs], trial order), build the
 * good/bad adaptive Parzen models and return log l(x) - log g(x) for the given candidates
 * (d_candidates, d_scores: float[n_problems][n_candidates]).  d_models (optional):
 * float[n_problems][2][cap+1][3] = (weight, mu, sigma) of the good then the bad model, cap = n_obs rounded up to 32.
 * d_lean_scores (optional): float[n_problems][n_candidates], the same candidates scored the way at2_tpe_sim_f32 scores them
 * inside a suggestion - without the weight normalisation and the truncation mass p_accept, which shift all candidates of a
 * problem alike: d_lean_scores - d_scores is constant per problem and the arg-max is the same.
 * d_samples (optional, needs d_lean_scores): float[n_problems][32], 32 draws per problem from the good model truncated to
 * [low, high), made the way a suggestion draws its candidates (tpe.GMM1's distribution, sampled without its rejection loop). */
int at2_tpe_score_f32(const float* d_obs, const float* d_loss, int32_t n_obs, int32_t n_problems, double low, double high,
                      const at2_tpe_config* tpe, const float* d_candidates, int32_t n_candidates, float* d_scores,
                      float* d_models, float* d_lean_scores, float* d_samples, void* stream);

/* ---- (d): closest-known-loss-function approximation ------------------------------------------------------------------ */

#define AT2_MAX_PARAMS 16

/* One hyperparameter of a problem domain (core/params.py:12-24).  For is_log the bounds are exponents of e. */
typedef struct at2_param {
    double min_val, max_val;
    double interval;        /* 0 = continuous; otherwise values are floor(v/interval)*interval (params.py:44-45) */
    double default_val;     /* init_val, used when the hyperparameter is not optimised (core/arm.py:29-42) */
    int32_t is_log;
    int32_t optimised;
} at2_param;

/* Domain in declaration order + the attribute order of an arm made by Arm().draw_hp_val (core/arm.py:44-55:
 * non-optimised defaults first, then the optimised ones), which is the summation order of the squared error. */
typedef struct at2_domain {
    int32_t n_params;
    int32_t _pad;
    at2_param p[AT2_MAX_PARAMS];
    int32_t order[AT2_MAX_PARAMS];
} at2_domain;

/* Batched nearest recorded arm: replaces the scan of benchmarks/known_loss_fn_problem.py:35-50 (+ Arm.normalize,
 * core/arm.py:57-70) for nq proposed arms at once.  d_db[A][D] and d_queries[nq][D] hold RAW hyperparameter values in
 * domain order.  Optional fused read-out (:61): d_loss[i] = d_curves[idx[i]][d_time[i] - 1].
 * float64, un-fused, first minimum wins -> indices are bit-identical to the reference's. */
int64_t at2_nn_workspace_bytes(int64_t nq, int32_t A, int32_t D);
int at2_nn_lookup_f64(const at2_domain* domain, const double* d_db, int32_t A, const double* d_queries, int64_t nq,
                      int32_t* d_idx, double* d_dist, const double* d_curves, int32_t L, const int32_t* d_time,
                      double* d_loss, void* d_workspace, int64_t workspace_bytes, void* stream);

/* Same, but query i ignores the recorded arms d_exclude[2i] <= a < d_exclude[2i+1]: the held-out fold of
 * experiments/known_functions_evaluation/cross_validate_approximation.py:57-66 without rebuilding the database. */
int at2_nn_lookup_excluding_f64(const at2_domain* domain, const double* d_db, int32_t A, const double* d_queries,
                                int64_t nq, const int32_t* d_exclude, int32_t* d_idx, double* d_dist, void* d_workspace,
                                int64_t workspace_bytes, void* stream);

/* Hyperband repetitions on the closest-known-function problem, fused: replaces the loop of
 * experiments/run_closest_loss_fn_approximation.py:183-194 for method sim(hb).  Per repetition: draw the arms
 * (core/arm.py:44-55) or take them from d_predrawn_arms[n_reps][n_arms][D], find each one's nearest recorded arm,
 * read its curve at the Hyperband checkpoints, run successive halving.  Outputs as at2_hb_outputs with real = double
 * (d_best_xy unused); optional d_nn_idx[n_reps][n_arms], d_best_arm_vals[n_reps][D]. */
int at2_hb_knownfn_f64(const at2_plan* plan, const at2_domain* domain, const double* d_db, int32_t A,
                       const double* d_curves, int32_t L, int32_t descending, int64_t n_reps, int64_t rep_offset,
                       uint64_t seed, const double* d_predrawn_arms, const at2_hb_outputs* out, int32_t* d_nn_idx,
                       double* d_best_arm_vals, void* stream);

/* ---- (e): EPDF-OFE kernel density estimate ----------------------------------------------------------------------- */

#define AT2_KDE_EPANECHNIKOV 0   /* the reference's kernel (plot_results.py:13) */
#define AT2_KDE_GAUSSIAN 1       /* extra */

/* Density of the n samples d_x on the G-point grid linspace(start, end, G) with the given bandwidth:
 * replaces sklearn's KernelDensity(kernel='epanechnikov', bandwidth).fit(x) + exp(score_samples(grid))
 * (experiments/simulation_evaluation/plot_results.py:11-17).  d_workspace: at2_kde_workspace_floats(n, G) floats - 0, and
 * d_workspace may then be NULL, for grids of up to 16 384 points: the CTAs' partial sums are accumulated as 64-bit
 * fixed-point integers in library-owned device memory (order-independent, so the density is deterministic). */
int64_t at2_kde_workspace_floats(int64_t n, int32_t G);
int at2_kde_f32(const float* d_x, int64_t n, double start, double end, int32_t G, double bandwidth, int32_t kernel,
                float* d_density, float* d_workspace, int64_t workspace_floats, void* stream);

/* The same estimate sharded over the ranks that hold the samples (N GPUs: every rank evaluates only its own slice of the
 * OFE vector).  at2_kde_partial_f32 writes the UNNORMALISED float64 kernel sums of this call's n samples on the G grid
 * points to elements slot_offset .. slot_offset + G - 1 of every target - the ranks' symmetric buffers mapped into this
 * process (peer stores), or one NVSwitch multicast address (is_multicast = 1: multimem.st) - rank r uses
 * slot_offset = r * G.  After a barrier across the ranks, at2_kde_combine_f64 adds the n_parts slots IN SLOT ORDER and
 * normalises by n_total samples: every rank obtains the same bits. */
int at2_kde_partial_f32(const float* d_x, int64_t n, double start, double end, int32_t G, double bandwidth, int32_t kernel,
                        double* const* d_targets, int32_t n_targets, int32_t is_multicast, int64_t slot_offset,
                        float* d_workspace, int64_t workspace_floats, void* stream);
int at2_kde_combine_f64(const double* d_parts, int32_t n_parts, int32_t G, int64_t n_total, double bandwidth, int32_t kernel,
                        float* d_density, void* stream);

/* Full-trajectory output (inspection / per-evaluator API, core/simulation_evaluator.py:84-116): raw trajectory
 * d_raw[a][0..R-1] of n_arms arms given their first value d_f1[a], target d_fn[a], family index and Philox stream id.
 * Arm `a` of repetition `rep` in at2_hb_sim_f32 has stream id (rep_offset + rep) * n_arms + a; for that id the values
 * at the Hyperband checkpoints are bit-identical to the fused kernel's. */
int at2_sim_trajectories_f32(const at2_plan* plan, const at2_sim_config* cfg, const void* d_tables, int64_t n_arms,
                             const float* d_f1, const float* d_fn, const int32_t* d_family, const uint64_t* d_stream_id,
                             uint64_t seed, float* d_raw, void* stream);

/* ---- post-processing statistics ("next" rows) ------------------------------------------------------------------------ */

/* Order profiles of a family of loss curves d_curves[n_curves][T] (experiments/simulation_evaluation/profiles.py):
 * d_dynamic[t-1], t = 1..T-1: mean over curves i of the fraction of other curves whose strict order relative to i is the
 * same at epochs t-1 and t (get_dynamic_order_profile, :29-53); d_ends[0]: the same between epochs 0 and T-1
 * (get_ends_order_profile, :56-81).  Either output may be NULL. */
int at2_order_profile_f32(const float* d_curves, int32_t n_curves, int32_t T, float* d_dynamic, float* d_ends, void* stream);

/* Two-sample Kolmogorov-Smirnov statistic of two SORTED samples (scipy.stats.ks_2samp's two-sided D, as used at
 * experiments/simulation_evaluation/plot_run_simulations_results.py:69-73). */
int at2_ks_statistic_f32(const float* d_a_sorted, int64_t n1, const float* d_b_sorted, int64_t n2, float* d_statistic,
                         void* stream);

/* Number of kernel launches the last at2_* call on this thread enqueued (for bench.py's gpu_launches). */
int32_t at2_last_launch_count(void);

/* Raw Philox4x32-10 block (known-answer tests): out[4] = philox(ctr[4], key[2]).  Host function. */
void at2_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);

/* FP32 FMA-only micro-benchmark used as the roofline denominator of the simulation kernel: launches a kernel doing
 * `iters` x 16 independent FFMAs per thread on `blocks` x `threads`; the caller times it with CUDA events and takes the
 * best of the operand-shape variants 0..2.  flops = 2 * iters * 16 * blocks * threads. */
int at2_fma_peak_probe(int32_t blocks, int32_t threads, int64_t iters, int32_t variant, float* d_sink, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* AT2_H */
