// Host-side pieces of libat2_b200: error state, Hyperband bracket plan, constant tables, Savitzky-Golay operator rows.
#include <cmath>
#include <cstring>
#include <vector>

#include "at2_common.cuh"

static thread_local char g_err[512] = "";
static thread_local int g_launches = 0;

void at2_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void at2_count_launch(int n) { g_launches += n; }
void at2_reset_launch_count() { g_launches = 0; }

static At2Debug g_debug = {};
const At2Debug& at2_debug() { return g_debug; }

extern "C" int at2_debug_option(const char* name, int32_t value) {
    AT2_REQUIRE(name != nullptr, "at2_debug_option: name is NULL");
    if (strcmp(name, "reset") == 0) {
        g_debug = At2Debug{};
        return AT2_OK;
    }
    const struct { const char* name; int* field; } table[] = {
        {"hb_reps_per_tile", &g_debug.hb_reps_per_tile}, {"hb_team_warps", &g_debug.hb_team_warps},
        {"hb_warps_per_cta", &g_debug.hb_warps_per_cta}, {"hb_no_split", &g_debug.hb_no_split},
        {"hb_equal_tiles", &g_debug.hb_equal_tiles}, {"hb_no_tile_halving", &g_debug.hb_no_tile_halving},
        {"hb_static_schedule", &g_debug.hb_static_schedule}, {"hb_singles", &g_debug.hb_singles},
        {"force_gtab", &g_debug.force_gtab}, {"lazy_reps_per_tile", &g_debug.lazy_reps_per_tile}, {"lazy_no_cache", &g_debug.lazy_no_cache},
        {"tpe_warps_per_cta", &g_debug.tpe_warps_per_cta}, {"kf_warps_per_cta", &g_debug.kf_warps_per_cta},
        {"kf_reps_per_tile", &g_debug.kf_reps_per_tile}};
    for (const auto& e : table)
        if (strcmp(name, e.name) == 0) {
            *e.field = value;
            return AT2_OK;
        }
    at2_set_error("at2_debug_option: unknown option '%s'", name);
    return AT2_EINVAL;
}

extern "C" const char* at2_version(void) { return "at2_b200 0.2 (sm_100a)"; }
extern "C" const char* at2_last_error(void) { return g_err; }
extern "C" int32_t at2_last_launch_count(void) { return g_launches; }

extern "C" void at2_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    At2U4 c = {ctr[0], ctr[1], ctr[2], ctr[3]};
    c = at2_philox10(c, key[0], key[1]);
    out[0] = c.x, out[1] = c.y, out[2] = c.z, out[3] = c.w;
}

// ---- Hyperband plan ---------------------------------------------------------------------------------------------------
// Mirrors the float arithmetic of optimisers/sequential/hyperband_optimiser.py:66-97 expression by expression
// (python int/int true division and int**negative-int both evaluate in IEEE double through libm pow).
static long long ipow_ll(long long b, int e) {
    long long r = 1;
    while (e-- > 0) r *= b;
    return r;
}

extern "C" int at2_bracket_plan(int32_t R, int32_t eta, at2_plan* plan) {
    AT2_REQUIRE(plan != nullptr, "at2_bracket_plan: plan is NULL");
    AT2_REQUIRE(R >= 2 && eta >= 2, "at2_bracket_plan: need R >= 2 and eta >= 2 (got R=%d eta=%d)", R, eta);
    memset(plan, 0, sizeof(*plan));
    int s_max = 0;  // floor(log_eta R), exact (:68-69 evaluates it with 64-digit mpmath to dodge the float-log bug)
    while (ipow_ll(eta, s_max + 1) <= (long long)R) ++s_max;
    const int s_min = s_max >= 2 ? 2 : 0;                       // :70
    const long long B = (long long)(s_max + 1) * R;             // :71
    plan->R = R;
    plan->eta = eta;
    int nb = 0, nr = 0, arms = 0, surv = 0;
    for (int s = s_max; s >= s_min; --s) {                      // :74
        if (nb >= AT2_MAX_BRACKETS) {
            at2_set_error("at2_bracket_plan: more than %d brackets", AT2_MAX_BRACKETS);
            return AT2_ELIMIT;
        }
        const double per = (double)B / (double)R / (double)(s + 1);
        const long long n = (long long)per * ipow_ll(eta, s);   // :75 int(ceil(int(B/R/(s+1)) * eta**s))
        const double r = s == 0 ? (double)R : (double)R * pow((double)eta, (double)-s);  // :76
        plan->bracket_n[nb] = (int32_t)n;
        plan->bracket_first_arm[nb] = arms;
        plan->bracket_first_round[nb] = nr;
        plan->bracket_r0[nb] = r;
        long long alive = n;
        for (int i = 0; i <= s; ++i) {                          // :89
            if (nr >= AT2_MAX_ROUNDS) {
                at2_set_error("at2_bracket_plan: more than %d rounds", AT2_MAX_ROUNDS);
                return AT2_ELIMIT;
            }
            const double n_i = i == 0 ? (double)n : (double)n * pow((double)eta, (double)-i);  // :90
            const double r_i = r * (double)ipow_ll(eta, i);                                     // :91
            const long long keep = (long long)(n_i / (double)eta);                              // :96
            plan->round_n_eval[nr] = (int32_t)alive;
            plan->round_keep[nr] = (int32_t)keep;
            // time = int(n_resources) (opt_function_simulation_problem.py:58); the value read is fs[time-1], and
            // python's negative indexing makes time == 0 read the LAST point (happens when R * eta**(-s) rounds
            // just below 1, e.g. R = 729, eta = 3 -> r = 0.9999999999999999): keep that quirk.
            const int time_i = (int)r_i;
            plan->round_time[nr] = time_i >= 1 ? time_i : R + time_i;
            plan->round_res[nr] = time_i;
            alive = keep < alive ? keep : alive;
            surv += (int)alive;
            ++nr;
        }
        arms += (int)n;
        ++nb;
    }
    plan->bracket_first_round[nb] = nr;
    plan->n_brackets = nb;
    plan->n_rounds = nr;
    plan->n_arms = arms;
    plan->n_surv = surv;
    // distinct checkpoint times, ascending
    int nc = 0;
    for (int i = 0; i < nr; ++i) {
        const int t = plan->round_time[i];
        AT2_REQUIRE(t >= 1 && t <= R, "at2_bracket_plan: round %d reads time %d outside [1,%d]", i, t, R);
        bool seen = false;
        for (int c = 0; c < nc; ++c) seen |= plan->ckpt_time[c] == t;
        if (!seen) {
            if (nc >= AT2_MAX_CKPTS) {
                at2_set_error("at2_bracket_plan: more than %d checkpoints", AT2_MAX_CKPTS);
                return AT2_ELIMIT;
            }
            plan->ckpt_time[nc++] = t;
        }
    }
    for (int a = 0; a < nc; ++a)
        for (int b = a + 1; b < nc; ++b)
            if (plan->ckpt_time[b] < plan->ckpt_time[a]) {
                int t = plan->ckpt_time[a];
                plan->ckpt_time[a] = plan->ckpt_time[b];
                plan->ckpt_time[b] = t;
            }
    plan->n_ckpts = nc;
    for (int i = 0; i < nr; ++i)
        for (int c = 0; c < nc; ++c)
            if (plan->ckpt_time[c] == plan->round_time[i]) plan->round_ckpt[i] = c;
    return AT2_OK;
}

// ---- Savitzky-Golay operator rows ---------------------------------------------------------------------------------------
// scipy.signal.savgol_filter(x, window, 3, mode='interp') is linear in x: output i is a least-squares cubic over a
// `window`-point stretch evaluated at i - the centred stretch for interior points, the first / last `window` points for
// the `window//2` outputs at either end (core/simulation_evaluator.py:76-82).
extern "C" int32_t at2_savgol_window(int32_t R) {
    int w = (int)(0.17 * (double)R + 6);  // :80
    return w % 2 == 0 ? w + 1 : w;        // :81
}

static void cubic_fit_weights(int window, double pos, double* w_out) {
    // weights w_j with  p(pos) = sum_j w_j x_j  for the LS cubic p through (j, x_j), j = 0..window-1
    const int half = window / 2;
    const long double sc = 1.0L / (long double)half;
    long double G[4][5];
    for (int a = 0; a < 4; ++a)
        for (int b = 0; b < 5; ++b) G[a][b] = 0;
    for (int j = 0; j < window; ++j) {
        const long double u = (long double)(j - half) * sc;
        long double pa = 1;
        for (int a = 0; a < 4; ++a) {
            long double pb = 1;
            for (int b = 0; b < 4; ++b) {
                G[a][b] += pa * pb;
                pb *= u;
            }
            pa *= u;
        }
    }
    const long double up = ((long double)pos - (long double)half) * sc;
    long double pp = 1;
    for (int a = 0; a < 4; ++a) {
        G[a][4] = pp;
        pp *= up;
    }
    for (int col = 0; col < 4; ++col) {  // Gaussian elimination, partial pivoting
        int piv = col;
        for (int r = col + 1; r < 4; ++r)
            if (fabsl(G[r][col]) > fabsl(G[piv][col])) piv = r;
        for (int b = 0; b < 5; ++b) {
            long double t = G[col][b];
            G[col][b] = G[piv][b];
            G[piv][b] = t;
        }
        for (int r = 0; r < 4; ++r) {
            if (r == col) continue;
            const long double f = G[r][col] / G[col][col];
            for (int b = col; b < 5; ++b) G[r][b] -= f * G[col][b];
        }
    }
    long double y[4];
    for (int a = 0; a < 4; ++a) y[a] = G[a][4] / G[a][a];
    for (int j = 0; j < window; ++j) {
        const long double u = (long double)(j - half) * sc;
        w_out[j] = (double)(y[0] + u * (y[1] + u * (y[2] + u * y[3])));
    }
}

extern "C" int at2_savgol_rows(int32_t R, int32_t n_rows, const int32_t* rows, double* out) {
    AT2_REQUIRE(rows && out && n_rows >= 0, "at2_savgol_rows: NULL argument");
    const int window = at2_savgol_window(R);
    // same failure the reference hits through scipy (e.g. R=6 -> window 7)
    AT2_REQUIRE(window <= R, "If mode is 'interp', window_length must be less than or equal to the size of x.");
    const int half = window / 2;
    std::vector<double> w(window);
    for (int i = 0; i < n_rows; ++i) {
        const int row = rows[i];
        AT2_REQUIRE(row >= 0 && row < R, "at2_savgol_rows: row %d outside [0,%d)", row, R);
        int first;
        double pos;
        if (row < half) {
            first = 0, pos = row;
        } else if (row >= R - half) {
            first = R - window, pos = row - first;
        } else {
            first = row - half, pos = half;
        }
        cubic_fit_weights(window, pos, w.data());
        for (int j = 0; j < R; ++j) out[(size_t)i * R + j] = 0.0;
        for (int j = 0; j < window; ++j) out[(size_t)i * R + first + j] = w[j];
    }
    return AT2_OK;
}

extern "C" int32_t at2_func_dims(int32_t func) {
    if (func >= AT2_FUNC_EGG && func <= AT2_FUNC_RASTRIGIN) return 2;
    return func == AT2_FUNC_EGG4 ? 4 : -1;
}

// ---- constant tables ----------------------------------------------------------------------------------------------------
static int check_cfg(const at2_plan* plan, const at2_sim_config* cfg) {
    AT2_REQUIRE(plan && cfg, "NULL plan/config");
    AT2_REQUIRE(plan->R >= 2 && plan->n_arms > 0 && plan->n_rounds > 0, "plan is empty (call at2_bracket_plan)");
    AT2_REQUIRE(cfg->n_families >= 1 && cfg->n_families <= AT2_MAX_FAMILIES, "n_families %d outside [1,%d]",
                cfg->n_families, AT2_MAX_FAMILIES);
    AT2_REQUIRE(at2_func_dims(cfg->func) > 0, "unknown func id %d", cfg->func);
    return AT2_OK;
}

extern "C" int64_t at2_hb_tables_bytes(const at2_plan* plan, const at2_sim_config* cfg, int32_t is_f64) {
    if (check_cfg(plan, cfg) != AT2_OK) return AT2_EINVAL;
    int slot[AT2_MAX_FAMILIES];
    At2TableLayout L{plan->R, at2_family_slot_map(cfg, slot), at2_cfg_has_smooth(cfg), at2_sg_stride(plan->n_ckpts)};
    return (int64_t)L.total() * (is_f64 ? 8 : 4);
}

template <typename T>
static int build_tables(const at2_plan* plan, const at2_sim_config* cfg, T* out) {
    const int R = plan->R;
    const bool f32 = sizeof(T) == 4;
    int slot[AT2_MAX_FAMILIES], first_of[AT2_MAX_FAMILIES];     // first_of[block] = a family that owns the block
    const int F = at2_family_slot_map(cfg, slot);              // blocks of records, one per distinct (ml_agg, nec_agg)
    for (int f = cfg->n_families - 1; f >= 0; --f) first_of[slot[f]] = f;
    At2TableLayout L{R, F, at2_cfg_has_smooth(cfg), at2_sg_stride(plan->n_ckpts)};
    for (int i = 0; i < L.total(); ++i) out[i] = (T)0;
    const double k = 2.0;  // simulation_evaluator.py:94
    const double log2e = 1.4426950408889634;
    T* rec = out + L.off_rec();
    T* pwt = out + L.off_pw();
    const int rows = L.rows();
    for (int t = 1; t < R; ++t) {
        const double m = (double)((R + 1) - t);                 // n - time with n = R + 1 (:98)
        const double root = sqrt(k * k + 4 * m);                // :37
        const double beta = (k + root) / (2 * m);               // :38
        const double alpha = k * beta + 1;                      // :39
        const double scale = 1 / beta;                          // :42
        const double d = alpha - 1.0 / 3.0;                     // Marsaglia & Tsang (2000)
        const double c = 1.0 / sqrt(9.0 * d);
        const double frac = (double)t / (double)(R - 1);
        bool is_ckpt = false;                                   // position t is read by some round of the plan
        for (int q = 0; q < plan->n_ckpts; ++q) is_ckpt |= plan->ckpt_time[q] - 1 == t;
        for (int f = 0; f < F; ++f) {
            T* row = rec + (size_t)AT2_REC * (f * rows + t);
            T* rpw = row + 4;
            const at2_family& fam = cfg->fam[first_of[f]];
            const double P = pow(frac, fam.nec_agg);          // :105
            const double P11 = pow(frac, 1.1 * fam.nec_agg);  // :112
            const double ml = fam.ml_agg / 100.0, w = d * scale;
            const double K1 = ml * (1.0 - P), K0 = P - 2.0 * ml * (1.0 - P);   // a (1-P) + P with a = (g-2) ml
            row[0] = (T)d;
            row[1] = (T)c;
            row[2] = (T)(d * log2e);
            if (f32) {
                row[3] = (T)(d * log2e + 32.0);
                rpw[0] = (T)(is_ckpt ? -w : w);
                rpw[1] = (T)(w * K1);
                rpw[2] = (T)K0;
                rpw[3] = (T)P11;
            } else {
                row[3] = (T)w;
                row[4] = (T)P;
                row[5] = (T)P11;
                row[6] = (T)(1.0 - P);                            // 1 - P, for the fused form of :103-107
            }
            T* prow = pwt + (size_t)AT2_PW * (f * rows + t);
            prow[0] = (T)K1, prow[1] = (T)K0, prow[2] = (T)P11;
        }
    }
    for (int f = 0; f < F; ++f) {   // sentinel row R: NaN slope -> the acceptance test can never pass
        T* row = rec + (size_t)AT2_REC * (f * rows + R);
        T* rpw = row + 4;
        row[0] = (T)1, row[1] = (T)NAN, row[2] = (T)log2e;
        if (f32)
            row[3] = (T)(log2e + 32.0), rpw[0] = (T)1, rpw[1] = (T)0, rpw[2] = (T)0, rpw[3] = (T)1;
        else
            row[3] = (T)1, row[4] = (T)1, row[5] = (T)1, row[6] = (T)0;
    }
    if (L.has_smooth) {
        int32_t rows[AT2_MAX_CKPTS];
        for (int c = 0; c < plan->n_ckpts; ++c) rows[c] = plan->ckpt_time[c] - 1;
        std::vector<double> W((size_t)plan->n_ckpts * R);
        const int rc = at2_savgol_rows(R, plan->n_ckpts, rows, W.data());
        if (rc != AT2_OK) return rc;
        T* sg = out + L.off_sg();
        for (int t = 0; t < R; ++t)
            for (int c = 0; c < plan->n_ckpts; ++c) sg[t * L.sg_stride + c] = (T)W[(size_t)c * R + t];
    }
    return AT2_OK;
}

extern "C" int32_t at2_family_slots(const at2_sim_config* cfg, int32_t* slots) {
    if (cfg == nullptr || slots == nullptr || cfg->n_families < 1 || cfg->n_families > AT2_MAX_FAMILIES) return AT2_EINVAL;
    int slot[AT2_MAX_FAMILIES];
    const int n = at2_family_slot_map(cfg, slot);
    for (int f = 0; f < cfg->n_families; ++f) slots[f] = slot[f];
    return n;
}

extern "C" int at2_hb_tables_build(const at2_plan* plan, const at2_sim_config* cfg, int32_t is_f64, void* host_out) {
    const int rc = check_cfg(plan, cfg);
    if (rc != AT2_OK) return rc;
    AT2_REQUIRE(host_out != nullptr, "at2_hb_tables_build: host_out is NULL");
    return is_f64 ? build_tables<double>(plan, cfg, (double*)host_out) : build_tables<float>(plan, cfg, (float*)host_out);
}
