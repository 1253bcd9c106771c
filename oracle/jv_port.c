/*
 * oracle/jv_port.c -- TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C restatement of the reference's seeded Jonker-Volgenant solve,
 * written from the semantics catalogued in SURVEY.md Appendix A.  It is the
 * checker the CUDA path is compared against (tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline leg are the only things allowed to load it) and
 * it is never linked into, imported by, or called from the product library.
 *
 * Parity status: PINNED.  tests/test_oracle_port.py checks this file against
 *   (a) the unmodified reference solver compiled into oracle/_ref/libreflap.so
 *       (oracle/Makefile) on seeded/noisy/tie-heavy instances, bit for bit, and
 *   (b) the reference's own known-answer vectors
 *       (LAP/lap/tests/test_lapjv.py:60-129, test_utils.py fixtures).
 *
 * Reference map (all relative to /root/reference):
 *   jvp_project          LAP/_lapjv_cpp/lapjv_seeded.cpp:38-48
 *   jvp_is_feasible      LAP/_lapjv_cpp/lapjv_seeded.cpp:9-17
 *   jvp_tighten_rows     LAP/_lapjv_cpp/lapjv_seeded.cpp:66-73
 *   jvp_greedy           LAP/_lapjv_cpp/lapjv_seeded.cpp:76-102
 *   jvp_count_tight      LAP/_lapjv_cpp/lapjv_seeded.cpp:105-113
 *   jvp_micro_arr        LAP/_lapjv_cpp/lapjv_seeded.cpp:136-159
 *   jvp_col_reduce       LAP/_lapjv_cpp/lapjv.cpp:8-72      (_ccrrt_dense)
 *   jvp_arr_pass         LAP/_lapjv_cpp/lapjv.cpp:76-149    (_carr_dense)
 *   jvp_level_collect    LAP/_lapjv_cpp/lapjv.cpp:153-171   (_find_dense)
 *   jvp_level_relax      LAP/_lapjv_cpp/lapjv.cpp:178-213   (_scan_dense)
 *   jvp_shortest_path    LAP/_lapjv_cpp/lapjv.cpp:221-282   (find_path_dense)
 *   jvp_augment_all      LAP/_lapjv_cpp/lapjv.cpp:286-319   (_ca_dense)
 *   jvp_cold_solve       LAP/_lapjv_cpp/lapjv.cpp:323-346   (lapjv_internal)
 *   jvp_lapjv_seeded     LAP/_lapjv_cpp/lapjv_seeded.cpp:19-173
 *
 * All arithmetic is IEEE binary64, evaluated left to right exactly as the
 * reference writes it; the file is compiled with -ffp-contract=off.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define JVP_BIG 1000000 /* LAP/_lapjv_cpp/lapjv.h:4 */

/* Counters filled in by every entry point (all optional: pass NULL). */
typedef struct jvp_trace {
    int64_t proj_triggers;   /* projection updates applied                   */
    int64_t tight_edges;     /* total_tight_edges                            */
    int64_t greedy_matched;  /* rows matched by the first-fit pass           */
    int64_t took_fallback;   /* 1 when the cold solve replaced the warm one  */
    int64_t micro_bumps;     /* micro-ARR column bumps                       */
    int64_t free_after_cr;   /* free rows after column reduction (cold)      */
    int64_t arr_iters;       /* augmenting-row-reduction iterations (cold)   */
    int64_t aug_paths;       /* shortest augmenting paths run                */
    int64_t collect_calls;   /* level-collect (_find_dense) calls            */
    int64_t relax_cols;      /* columns relaxed from (one per SCAN column)   */
    int64_t rc;              /* return code                                  */
    /* optional per-path log: (start row, end column) pairs, capacity given */
    int32_t *path_log;
    int64_t path_log_cap;
} jvp_trace;

static void trace_zero(jvp_trace *t)
{
    if (!t) return;
    int32_t *log = t->path_log;
    int64_t cap = t->path_log_cap;
    memset(t, 0, sizeof(*t));
    t->path_log = log;
    t->path_log_cap = cap;
}

/* ------------------------------------------------------------------------ */
/* warm-start front end                                                      */
/* ------------------------------------------------------------------------ */

/* Row-major Gauss-Seidel clamp: whenever u_i + v_j overshoots c_ij by more
 * than eps, both potentials give up half of the overshoot, in place. */
static int64_t jvp_project(const double *c, int n, double *u, double *v, double eps)
{
    int64_t fired = 0;
    for (int r = 0; r < n; ++r) {
        const double *row = c + (size_t)r * (size_t)n;
        for (int k = 0; k < n; ++k) {
            double over = u[r] + v[k] - row[k];
            if (over > eps) {
                double half = over / 2.0;
                u[r] -= half;
                v[k] -= half;
                ++fired;
            }
        }
    }
    return fired;
}

static int jvp_is_feasible(const double *c, int n, const double *u, const double *v, double eps)
{
    for (int r = 0; r < n; ++r) {
        const double *row = c + (size_t)r * (size_t)n;
        for (int k = 0; k < n; ++k)
            if (row[k] - u[r] - v[k] < -eps) return 0;
    }
    return 1;
}

/* u_r := min_k (c_rk - v_k); the running minimum keeps the earlier value on
 * ties/unordered compares, like std::min(acc, candidate). */
static void jvp_tighten_rows(const double *c, int n, double *u, const double *v)
{
    for (int r = 0; r < n; ++r) {
        const double *row = c + (size_t)r * (size_t)n;
        double best = INFINITY;
        for (int k = 0; k < n; ++k) {
            double red = row[k] - v[k];
            if (red < best) best = red;
        }
        u[r] = best;
    }
}

static int jvp_edge_is_tight(double cij, double ui, double vj, double tol)
{
    return fabs(cij - ui - vj) <= tol;
}

/* First-fit: each row, in order, grabs its lowest-index tight column that no
 * earlier row took. */
static int64_t jvp_greedy(const double *c, int n, const double *u, const double *v,
                          double tol, int *x, int *y)
{
    int64_t matched = 0;
    unsigned char *taken = (unsigned char *)calloc((size_t)n, 1);
    if (!taken) return -1;
    for (int r = 0; r < n; ++r) {
        const double *row = c + (size_t)r * (size_t)n;
        for (int k = 0; k < n; ++k) {
            if (taken[k]) continue;
            if (jvp_edge_is_tight(row[k], u[r], v[k], tol)) {
                x[r] = k;
                y[k] = r;
                taken[k] = 1;
                ++matched;
                break;
            }
        }
    }
    free(taken);
    return matched;
}

static int64_t jvp_count_tight(const double *c, int n, const double *u, const double *v, double tol)
{
    int64_t cnt = 0;
    for (int r = 0; r < n; ++r) {
        const double *row = c + (size_t)r * (size_t)n;
        for (int k = 0; k < n; ++k) cnt += jvp_edge_is_tight(row[k], u[r], v[k], tol);
    }
    return cnt;
}

/* For every still-free row: if its best reduced cost is separated from the
 * runner-up by more than tol and the best column is unmatched, raise that
 * column's potential by the gap. */
static int64_t jvp_micro_arr(const double *c, int n, const double *u, double *v, double tol,
                             const int *free_rows, int n_free, const int *y_after_greedy)
{
    int64_t bumps = 0;
    for (int f = 0; f < n_free; ++f) {
        int r = free_rows[f];
        const double *row = c + (size_t)r * (size_t)n;
        double lo1 = INFINITY, lo2 = INFINITY;
        int at = -1;
        for (int k = 0; k < n; ++k) {
            double red = row[k] - u[r] - v[k];
            if (red < lo1) {
                lo2 = lo1;
                lo1 = red;
                at = k;
            } else if (red < lo2) {
                lo2 = red;
            }
        }
        /* "at is in free_cols" == column was unmatched after the greedy pass */
        if (at >= 0 && lo2 - lo1 > tol && y_after_greedy[at] < 0) {
            v[at] += lo2 - lo1;
            ++bumps;
        }
    }
    return bumps;
}

/* ------------------------------------------------------------------------ */
/* shortest augmenting paths                                                  */
/* ------------------------------------------------------------------------ */

/* Partition order[] so that the columns of minimum dist among order[lo..n)
 * sit in order[lo..hi).  Every prefix-minimum record AND every tie swaps, so
 * the resulting order is history dependent -- that is part of the contract. */
static int jvp_level_collect(int n, int lo, const double *dist, int *order)
{
    int hi = lo + 1;
    double level = dist[order[lo]];
    for (int k = lo + 1; k < n; ++k) {
        int col = order[k];
        double dk = dist[col];
        if (!(dk <= level)) continue;
        if (dk < level) {
            level = dk;
            hi = lo;
        }
        order[k] = order[hi];
        order[hi] = col;
        ++hi;
    }
    return hi;
}

/* Relax all TODO columns from each SCAN column in turn.  Returns an unmatched
 * column reaching the current level (first in order[] position), or -1 after
 * the SCAN list drained (then lo/hi are written back; on early return they
 * are deliberately left untouched, as in the reference). */
static int jvp_level_relax(const double *c, int n, int *plo, int *phi, double *dist,
                           int *order, int *pred, const int *y, const double *v,
                           jvp_trace *t)
{
    int lo = *plo, hi = *phi;
    while (lo != hi) {
        int from_col = order[lo++];
        int via_row = y[from_col];
        const double *row = c + (size_t)via_row * (size_t)n;
        double level = dist[from_col];
        double slack = row[from_col] - v[from_col] - level;
        if (t) t->relax_cols++;
        for (int k = hi; k < n; ++k) {
            int col = order[k];
            double cand = row[col] - v[col] - slack;
            if (cand < dist[col]) {
                dist[col] = cand;
                pred[col] = via_row;
                if (cand == level) {
                    if (y[col] < 0) return col;
                    order[k] = order[hi];
                    order[hi] = col;
                    ++hi;
                }
            }
        }
    }
    *plo = lo;
    *phi = hi;
    return -1;
}

static int jvp_shortest_path(const double *c, int n, int start_row, const int *y, double *v,
                             int *pred, double *dist, int *order, jvp_trace *t)
{
    const double *row0 = c + (size_t)start_row * (size_t)n;
    for (int k = 0; k < n; ++k) {
        order[k] = k;
        pred[k] = start_row;
        dist[k] = row0[k] - v[k];
    }
    int lo = 0, hi = 0, n_done = 0, sink = -1;
    while (sink < 0) {
        if (lo == hi) {
            n_done = lo;
            hi = jvp_level_collect(n, lo, dist, order);
            if (t) t->collect_calls++;
            for (int k = lo; k < hi; ++k)
                if (y[order[k]] < 0) sink = order[k]; /* last one wins */
        }
        if (sink < 0) sink = jvp_level_relax(c, n, &lo, &hi, dist, order, pred, y, v, t);
    }
    double level = dist[order[lo]];
    for (int k = 0; k < n_done; ++k) {
        int col = order[k];
        v[col] += dist[col] - level;
    }
    return sink;
}

static int jvp_augment_all(const double *c, int n, int n_free, const int *free_rows,
                           int *x, int *y, double *v, jvp_trace *t)
{
    int *pred = (int *)malloc(sizeof(int) * (size_t)n);
    int *order = (int *)malloc(sizeof(int) * (size_t)n);
    double *dist = (double *)malloc(sizeof(double) * (size_t)n);
    if (!pred || !order || !dist) {
        free(pred);
        free(order);
        free(dist);
        return -1;
    }
    for (int f = 0; f < n_free; ++f) {
        int root = free_rows[f];
        int col = jvp_shortest_path(c, n, root, y, v, pred, dist, order, t);
        if (t) {
            if (t->path_log && t->aug_paths < t->path_log_cap) {
                t->path_log[2 * t->aug_paths] = root;
                t->path_log[2 * t->aug_paths + 1] = col;
            }
            t->aug_paths++;
        }
        int r;
        do {
            r = pred[col];
            y[col] = r;
            int prev = x[r];
            x[r] = col;
            col = prev;
        } while (r != root);
    }
    free(pred);
    free(order);
    free(dist);
    return 0;
}

/* ------------------------------------------------------------------------ */
/* cold solve (the fallback)                                                  */
/* ------------------------------------------------------------------------ */

static int jvp_col_reduce(const double *c, int n, int *free_rows, int *x, int *y, double *v)
{
    for (int k = 0; k < n; ++k) {
        x[k] = -1;
        v[k] = JVP_BIG;
        y[k] = 0;
    }
    for (int r = 0; r < n; ++r) {
        const double *row = c + (size_t)r * (size_t)n;
        for (int k = 0; k < n; ++k)
            if (row[k] < v[k]) {
                v[k] = row[k];
                y[k] = r;
            }
    }
    unsigned char *sole = (unsigned char *)malloc((size_t)n);
    if (!sole) return -1;
    memset(sole, 1, (size_t)n);
    for (int k = n - 1; k >= 0; --k) {
        int r = y[k];
        if (x[r] < 0) {
            x[r] = k;
        } else {
            sole[r] = 0;
            y[k] = -1;
        }
    }
    int n_free = 0;
    for (int r = 0; r < n; ++r) {
        if (x[r] < 0) {
            free_rows[n_free++] = r;
        } else if (sole[r]) {
            const double *row = c + (size_t)r * (size_t)n;
            int own = x[r];
            double m = JVP_BIG;
            for (int k = 0; k < n; ++k) {
                if (k == own) continue;
                double red = row[k] - v[k];
                if (red < m) m = red;
            }
            v[own] -= m;
        }
    }
    free(sole);
    return n_free;
}

static int jvp_arr_pass(const double *c, int n, int n_free, int *free_rows, int *x, int *y,
                        double *v, jvp_trace *t)
{
    unsigned int cursor = 0, steps = 0;
    int deferred = 0;
    while (cursor < (unsigned int)n_free) {
        ++steps;
        if (t) t->arr_iters++;
        int r = free_rows[cursor++];
        const double *row = c + (size_t)r * (size_t)n;
        int k1 = 0, k2 = -1;
        double b1 = row[0] - v[0], b2 = JVP_BIG;
        for (int k = 1; k < n; ++k) {
            double red = row[k] - v[k];
            if (red < b2) {
                if (red >= b1) {
                    b2 = red;
                    k2 = k;
                } else {
                    b2 = b1;
                    b1 = red;
                    k2 = k1;
                    k1 = k;
                }
            }
        }
        int owner = y[k1];
        double lowered = v[k1] - (b2 - b1);
        int does_lower = lowered < v[k1];
        if (steps < cursor * (unsigned int)n) {
            if (does_lower) {
                v[k1] = lowered;
            } else if (owner >= 0 && k2 >= 0) {
                k1 = k2;
                owner = y[k2];
            }
            if (owner >= 0) {
                if (does_lower)
                    free_rows[--cursor] = owner;
                else
                    free_rows[deferred++] = owner;
            }
        } else if (owner >= 0) {
            free_rows[deferred++] = owner;
        }
        x[r] = k1;
        y[k1] = r;
    }
    return deferred;
}

static int jvp_cold_solve(const double *c, int n, int *x, int *y, jvp_trace *t, double *v_out)
{
    int *free_rows = (int *)malloc(sizeof(int) * (size_t)n);
    double *v = (double *)malloc(sizeof(double) * (size_t)n);
    if (!free_rows || !v) {
        free(free_rows);
        free(v);
        return -1;
    }
    int left = jvp_col_reduce(c, n, free_rows, x, y, v);
    if (t && left >= 0) t->free_after_cr = left;
    for (int pass = 0; left > 0 && pass < 2; ++pass) left = jvp_arr_pass(c, n, left, free_rows, x, y, v, t);
    if (left > 0) left = jvp_augment_all(c, n, left, free_rows, x, y, v, t);
    if (v_out) memcpy(v_out, v, sizeof(double) * (size_t)n);
    free(v);
    free(free_rows);
    return left;
}

/* ------------------------------------------------------------------------ */
/* exported entry points                                                      */
/* ------------------------------------------------------------------------ */

int jvp_lapjv(const double *c, int n, int *x, int *y, jvp_trace *t)
{
    trace_zero(t);
    if (n <= 0) return -2;
    int rc = jvp_cold_solve(c, n, x, y, t, NULL);
    if (t) t->rc = rc;
    return rc;
}

/* Cold solve that also hands back the final column potentials (optimal duals:
 * u_i = c[i][x_i] - v[x_i]).  Test/bench helper for building oracle-dual seeds. */
int jvp_lapjv_duals(const double *c, int n, int *x, int *y, double *v_out)
{
    if (n <= 0) return -2;
    return jvp_cold_solve(c, n, x, y, NULL, v_out);
}

/* Front end only: projected/tightened potentials, greedy matching, tight
 * count.  Used by the tests to localise a GPU/CPU divergence. */
static int front_end_impl(const double *c, int n, const double *u_seed, const double *v_seed,
                          double eps, double *u, double *v, int *x, int *y, jvp_trace *t,
                          int64_t *tight_out)
{
    if (n <= 0) return -2;
    memcpy(u, u_seed, sizeof(double) * (size_t)n);
    memcpy(v, v_seed, sizeof(double) * (size_t)n);
    for (int k = 0; k < n; ++k) x[k] = y[k] = -1;
    int64_t fired = jvp_project(c, n, u, v, eps);
    if (!jvp_is_feasible(c, n, u, v, eps)) return -3;
    jvp_tighten_rows(c, n, u, v);
    double tol = eps > 1e-9 ? eps : 1e-9;
    int64_t m = jvp_greedy(c, n, u, v, tol, x, y);
    if (m < 0) return -1;
    int64_t tight = jvp_count_tight(c, n, u, v, tol);
    if (tight_out) *tight_out = tight;
    if (t) {
        t->proj_triggers = fired;
        t->greedy_matched = m;
        t->tight_edges = tight;
        t->took_fallback = (double)tight < 1.2 * n;
    }
    return 0;
}

int jvp_front_end(const double *c, int n, const double *u_seed, const double *v_seed, double eps,
                  double *u, double *v, int *x, int *y, jvp_trace *t)
{
    trace_zero(t);
    int rc = front_end_impl(c, n, u_seed, v_seed, eps, u, v, x, y, t, NULL);
    if (t) t->rc = rc;
    return rc;
}

int jvp_lapjv_seeded(const double *c, int n_rows, int n_cols, long long *x_out, long long *y_out,
                     const double *u_seed, const double *v_seed, double eps, jvp_trace *t)
{
    trace_zero(t);
    if (n_rows <= 0 || n_cols <= 0) return -2;
    if (n_rows != n_cols) return -4;
    const int n = n_rows;
    int rc = 0;
    double *u = (double *)malloc(sizeof(double) * (size_t)n);
    double *v = (double *)malloc(sizeof(double) * (size_t)n);
    int *x = (int *)malloc(sizeof(int) * (size_t)n);
    int *y = (int *)malloc(sizeof(int) * (size_t)n);
    int *y0 = (int *)malloc(sizeof(int) * (size_t)n);
    int *free_rows = (int *)malloc(sizeof(int) * (size_t)n);
    if (!u || !v || !x || !y || !y0 || !free_rows) {
        rc = -1;
        goto done;
    }
    int64_t tight = 0;
    rc = front_end_impl(c, n, u_seed, v_seed, eps, u, v, x, y, t, &tight);
    if (rc != 0) goto done;
    {
        double tol = eps > 1e-9 ? eps : 1e-9;
        int n_free = 0;
        for (int r = 0; r < n; ++r)
            if (x[r] < 0) free_rows[n_free++] = r;
        if ((double)tight < 1.2 * n) {
            /* warm start judged useless: cold solve on the raw matrix */
            rc = jvp_cold_solve(c, n, x, y, t, NULL);
            if (t) t->took_fallback = 1;
            if (rc != 0) goto done;
        } else if (n_free > 0) {
            memcpy(y0, y, sizeof(int) * (size_t)n);
            int64_t b = jvp_micro_arr(c, n, u, v, tol, free_rows, n_free, y0);
            if (t) t->micro_bumps = b;
            rc = jvp_augment_all(c, n, n_free, free_rows, x, y, v, t);
            if (rc != 0) goto done;
        }
        for (int k = 0; k < n; ++k) {
            x_out[k] = x[k];
            y_out[k] = y[k];
        }
    }
done:
    if (t) t->rc = rc;
    free(u);
    free(v);
    free(x);
    free(y);
    free(y0);
    free(free_rows);
    return rc;
}
