/*
 * spx_oracle.c — CPU restatement of the reference's tableau pivot loop.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under simplex_method_solver_b200/ may
 * import, link or execute this file; only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs use it, and only as the
 * checker / the reported CPU baseline.
 *
 * Parity status: PINNED.  The reference ships no tests or golden vectors
 * (SURVEY.md §4), but its solver is stdlib-only Python, so it was executed in
 * the build container and its outputs are committed as fixtures under
 * tests/golden/ (generator: tests/golden/make_golden.py).  This restatement is
 * checked bit-for-bit against those fixtures by tests/test_oracle.py, and —
 * when /root/reference is present — against the live reference on random LPs.
 *
 * What is restated (all citations are into /root/reference/src/simplex.py):
 *   orc_pick    : pick_element()        simplex.py:70-141
 *   orc_update  : recalculate_matrix()  simplex.py:149-177 (arithmetic only)
 *   orc_solve   : the driving loop      simplex.py:179-199 / :261-269
 *   orc_extract : find_optimum(), f()   simplex.py:48-68  (generalised to m vars)
 *
 * Extension rule (NOT reference behaviour): orc_pick_rule / orc_solve_rule with
 * rule = ORC_RULE_DANTZIG replace ONLY the entering-column choice of simplex.py:94-98 (first negative
 * f cell) by "the most negative f cell, lowest index on ties" — the rule BASELINE.json's north star
 * words; phase 1 (:72-91), the ratio scan (:107-136) and the pivot arithmetic are the reference's.
 * There is no reference implementation to pin it against: tests/test_oracle.py checks it against an
 * independent pure-Python statement of the same rule built on the reference's own ratio scan.
 *
 * Data layout ("reference flat"): the reference's ragged list-of-lists
 * flattened row-major: rows 0..n-1 have m+1 cells [a_1..a_m, b], then the
 * f row with exactly m cells (simplex.py:36-39).  cells = n*(m+1)+m.
 *
 * Arithmetic: IEEE-754 binary64, every product, difference and quotient
 * rounded separately, exactly as CPython floats do.  Build with
 * -ffp-contract=off and without -ffast-math (see oracle/Makefile).
 *
 * Labels: the reference keeps strings 'x1'..'xm','-b' / 'y1'..'yn','f'
 * (simplex.py:30-33).  Here a label is an int32 code: x_j -> j-1 (0..m-1),
 * y_i -> m+i-1 (m..m+n-1).  rowlab has m entries (the header over the
 * columns, '-b' is implicit), collab has n entries ('f' implicit).  The swap
 * of simplex.py:152 is rowlab[c] <-> collab[r].
 */
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#include <stdlib.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_PIVOT      1   /* pick_element returned (True, r, c, e)          */
#define ORC_OPTIMAL    0   /* pick_element returned (False, x1, x2, f)       */
#define ORC_INCORRECT -1   /* ValueError("incorrect system")                 */
#define ORC_NOCONV    -2   /* ValueError("simplex method does not converge") */
#define ORC_CAP       -3   /* max_pivots reached (the reference has no cap)  */

static inline size_t row_off(int i, int m) { return (size_t)i * (size_t)(m + 1); }

#define ORC_RULE_REFERENCE 0   /* first negative f cell, simplex.py:94-98 */
#define ORC_RULE_DANTZIG   1   /* most negative f cell, lowest index on ties (extension) */

/* simplex.py:70-141, statement by statement; `rule` only changes the entering column (:94-98). */
int orc_pick_rule(const double *T, int n, int m, int rule, int *r_out, int *c_out, double *e_out)
{
    int target_row = -1, target_col = -1;

    /* :72-76  first row whose '-b' cell is negative */
    for (int i = 0; i < n; ++i)
        if (T[row_off(i, m) + m] < 0) { target_row = i; break; }

    if (target_row >= 0) {
        /* :81-85  first strictly positive cell in that row */
        const double *row = T + row_off(target_row, m);
        for (int j = 0; j < m; ++j)
            if (row[j] > 0) { target_col = j; break; }
        if (target_col < 0) return ORC_INCORRECT;               /* :88-89 */
        *r_out = target_row; *c_out = target_col; *e_out = row[target_col];
        return ORC_PIVOT;                                        /* :91 */
    }

    /* :94-98  first negative cell of the f row (m cells, no constant) */
    const double *f = T + row_off(n, m);
    if (rule == ORC_RULE_DANTZIG) {
        /* extension: the most negative cell; '<' keeps the lowest index on ties; NaN never enters */
        for (int j = 0; j < m; ++j)
            if (f[j] < 0 && (target_col < 0 || f[j] < f[target_col])) target_col = j;
    } else {
        for (int j = 0; j < m; ++j)
            if (f[j] < 0) { target_col = j; break; }
    }
    if (target_col < 0) return ORC_OPTIMAL;                      /* :101-103 */

    /* :107-136  the sequential ratio scan, kept as the state machine it is */
    int first_try = 1;
    double min_val = 1;
    for (int i = 0; i < n; ++i) {
        double a = T[row_off(i, m) + target_col];
        if (a == 0) continue;                                    /* :112 */
        double val = T[row_off(i, m) + m] / a;                   /* :115 */
        if (first_try) { min_val = val; target_row = i; first_try = 0; continue; }
        if (val == 0 && min_val > 0) { min_val = val; target_row = i; continue; }    /* :123 */
        if (val < 0 && 0 <= min_val) { min_val = val; target_row = i; continue; }    /* :128 */
        if (min_val <= val && val < 0) { min_val = val; target_row = i; continue; }  /* :133 */
    }
    if (first_try || min_val > 0) return ORC_NOCONV;             /* :138-139 */
    *r_out = target_row; *c_out = target_col;
    *e_out = T[row_off(target_row, m) + target_col];
    return ORC_PIVOT;                                            /* :141 */
}

int orc_pick(const double *T, int n, int m, int *r_out, int *c_out, double *e_out)
{
    return orc_pick_rule(T, n, m, ORC_RULE_REFERENCE, r_out, c_out, e_out);
}

/* simplex.py:155-175: out of place, every read from the old table. */
void orc_update(const double *T, double *N, int n, int m, int r, int c)
{
    const double p = T[row_off(r, m) + c];
    const double *prow = T + row_off(r, m);
#pragma omp parallel for schedule(static) if ((size_t)n * (size_t)m > 65536)
    for (int i = 0; i <= n; ++i) {
        const int w = (i < n) ? m + 1 : m;       /* the f row has m cells */
        const double *src = T + row_off(i, m);
        double *dst = N + row_off(i, m);
        if (i == r) {
            for (int j = 0; j < w; ++j) dst[j] = -src[j] / p;    /* :155-156 */
            dst[c] = 1.0 / p;                                    /* :163 */
        } else {
            const double ci = src[c];
            for (int j = 0; j < w; ++j)
                dst[j] = (src[j] * p - prow[j] * ci) / p;        /* :173-175 */
            dst[c] = ci / p;                                     /* :159-160 */
        }
    }
}

/* simplex.py:152 */
static inline void swap_labels(int32_t *rowlab, int32_t *collab, int r, int c)
{
    int32_t t = rowlab[c]; rowlab[c] = collab[r]; collab[r] = t;
}

void orc_init_labels(int32_t *rowlab, int32_t *collab, int n, int m)
{
    for (int j = 0; j < m; ++j) rowlab[j] = j;          /* 'x1'..'xm'  :30 */
    for (int i = 0; i < n; ++i) collab[i] = m + i;      /* 'y1'..'yn'  :31 */
}

/* find_optimum (:51-68) for every x_j, and c.x with the ORIGINAL function
 * (:48-49 uses self.function, which is never mutated).  x has m entries. */
void orc_extract(const double *T, int n, int m, const int32_t *collab,
                 const double *function, double *x, double *obj2, double *objm)
{
    for (int j = 0; j < m; ++j) x[j] = 0.0;
    /* the reference takes the FIRST row carrying the label (index_of, :52-56);
     * labels are unique, so any order gives the same answer. */
    for (int i = 0; i < n; ++i)
        if (collab[i] >= 0 && collab[i] < m) x[collab[i]] = T[row_off(i, m) + m];
    if (obj2) *obj2 = (m >= 2) ? function[0] * x[0] + function[1] * x[1] : 0.0;   /* :49 */
    if (objm) {
        double s = 0.0;
        for (int j = 0; j < m; ++j) s += function[j] * x[j];
        *objm = s;
    }
}

/*
 * The loop of simplex.py:261-269 (pick -> pivot), with a cap.
 *   T        in/out, reference-flat, cells doubles; holds the final table
 *   scratch  cells doubles (ping-pong partner)
 *   trace    optional [max_pivots][2] int32 (r,c) per pivot
 *   snaps    optional [(max_pivots+1)][cells] doubles: table BEFORE pivot k
 *            at slot k, final table at slot npiv (the Info.table sequence,
 *            simplex.py:181,198)
 * returns the final status; *npiv_out = pivots done.
 */
int orc_solve_rule(double *T, double *scratch, int n, int m, int rule, int64_t max_pivots,
                   int32_t *trace, double *snaps, int32_t *rowlab, int32_t *collab,
                   int64_t *npiv_out)
{
    const size_t cells = (size_t)n * (size_t)(m + 1) + (size_t)m;
    double *cur = T, *nxt = scratch;
    int64_t k = 0;
    int status;
    if (snaps) memcpy(snaps, cur, cells * sizeof(double));
    for (;;) {
        int r = 0, c = 0; double e = 0;
        status = orc_pick_rule(cur, n, m, rule, &r, &c, &e);
        if (status != ORC_PIVOT) break;
        if (k >= max_pivots) { status = ORC_CAP; break; }
        if (trace) { trace[2 * k] = r; trace[2 * k + 1] = c; }
        orc_update(cur, nxt, n, m, r, c);
        if (rowlab && collab) swap_labels(rowlab, collab, r, c);
        double *t = cur; cur = nxt; nxt = t;
        ++k;
        if (snaps) memcpy(snaps + (size_t)k * cells, cur, cells * sizeof(double));
    }
    if (cur != T) memcpy(T, cur, cells * sizeof(double));
    *npiv_out = k;
    return status;
}

int orc_solve(double *T, double *scratch, int n, int m, int64_t max_pivots,
              int32_t *trace, double *snaps, int32_t *rowlab, int32_t *collab,
              int64_t *npiv_out)
{
    return orc_solve_rule(T, scratch, n, m, ORC_RULE_REFERENCE, max_pivots, trace, snaps, rowlab, collab, npiv_out);
}

/*
 * Independent LPs of one shape, LP k at T + k*cells.  trace is
 * [B][max_pivots][2] (optional), labels [B][m] / [B][n] (optional, must be
 * initialised by the caller or NULL), x is [B][m], obj2 [B], status [B],
 * npiv [B].  function is [B][m] = the original f rows (copied by the caller
 * before the solve, since T is updated in place).
 */
void orc_solve_batched(double *T, int64_t B, int n, int m, int64_t max_pivots,
                       int32_t *trace, int32_t *rowlab, int32_t *collab,
                       const double *function, double *x, double *obj2,
                       int32_t *status, int32_t *npiv)
{
    const size_t cells = (size_t)n * (size_t)(m + 1) + (size_t)m;
#pragma omp parallel
    {
        double *scratch = (double *)malloc(cells * sizeof(double));
        int32_t *rl = (int32_t *)malloc((size_t)(m > 0 ? m : 1) * sizeof(int32_t));
        int32_t *cl = (int32_t *)malloc((size_t)(n > 0 ? n : 1) * sizeof(int32_t));
#pragma omp for schedule(dynamic, 256)
        for (int64_t k = 0; k < B; ++k) {
            int32_t *rlk = rowlab ? rowlab + k * m : rl;
            int32_t *clk = collab ? collab + k * n : cl;
            orc_init_labels(rlk, clk, n, m);
            int64_t np = 0;
            int st = orc_solve(T + k * cells, scratch, n, m, max_pivots,
                               trace ? trace + k * max_pivots * 2 : NULL, NULL,
                               rlk, clk, &np);
            status[k] = st;
            npiv[k] = (int32_t)np;
            orc_extract(T + k * cells, n, m, clk, function + k * m,
                        x + k * m, obj2 ? obj2 + k : NULL, NULL);
        }
        free(scratch); free(rl); free(cl);
    }
}

int orc_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void orc_set_num_threads(int t)
{
#ifdef _OPENMP
    if (t > 0) omp_set_num_threads(t);
#else
    (void)t;
#endif
}
