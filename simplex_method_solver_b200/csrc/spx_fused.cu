// spx_fused.cu — K6: F pivots per pass over the tableau (temporal blocking of the rank-1 updates).
//
// recalculate_matrix() (/root/reference/src/simplex.py:143-177) touches every cell once per pivot:
// 16 B of HBM traffic per cell per pivot is the roofline of a pivot-at-a-time implementation.  But
// what pick_element() (:70-141) needs to choose a pivot is only O(n + m) cells of the current
// table — the b column, one row (f, or the phase-1 row), one column — plus, to apply it, the pivot
// row and column.  Any single cell of the table "after i more pivots" can be evaluated lazily from
// the stored table by replaying those i rank-1 updates on that one cell, with exactly the
// reference's operation order and roundings (:156, :160, :163, :173-175).  So:
//
//   block_price_kernel  (one CTA) chooses the next F pivots from the MATERIALISED table k without
//       touching the body: for level i it evaluates the needed row/column of the virtual table
//       k+i through the i pending levels, runs the reference's selection rules on them, and records
//       the level: (r_i, c_i, p_i), ROW_i = pivot row and COL_i = pivot column of table k+i.
//   update_fused_kernel  streams the body ONCE and applies all F levels to every cell in registers:
//       16 B of HBM traffic per cell per F pivots.  Same tiling as update_tiled_kernel; per level
//       the 512-double slice of ROW_i and the slice of COL_i are staged in shared memory by TMA.
//
// Every cell still goes through the same sequence of separately rounded operations as in the
// pivot-at-a-time path, so the pivot sequence and every bit of the table are unchanged
// (tests/test_gpu_parity.py runs this loop against the same goldens and the oracle).
#include <cooperative_groups.h>

#include <new>

#include "spx_block.cuh"

namespace cg = cooperative_groups;

namespace {

using namespace spx;

constexpr int FUSE_MAX       = 8;      // levels per pass at most
constexpr int PRICE_THREADS  = 1024;
constexpr int FUP_THREADS    = 256;
constexpr int FUP_TC         = 2 * FUP_THREADS;
constexpr int FUP_TR         = 32;     // rows per tile: the per-level row slices are amortised over 4 batches
constexpr int FUP_UNROLL     = 8;
constexpr int PRICE_BATCH    = 8;
constexpr int COOP_THREADS   = 512;    // per CTA of the cooperative pricing kernels: fewer CTAs = cheaper grid barriers

struct Level { int32_t r; int32_t c; double p; };
struct alignas(128) PlanHeader {
    int32_t f;          // levels to apply in this pass (0: nothing to do)
    int32_t src;        // index (0/1) of the ping-pong buffer that holds the input table
    int32_t pad[2];
    Level   lvl[FUSE_MAX];
    int32_t owner[FUSE_MAX];   // column-sharded: the rank whose plane holds COL_l (0 on one GPU)
};

struct PriceArgs {
    double *A[2];
    double *b[2];
    int n, m;
    int64_t ld, cbd;        // leading dimension of the body / stride of the COLS planes
    int rule, F;
    spx_state *st;
    PlanHeader *plan;
    double *ROWS;           // [FUSE_MAX][ld]
    double *COLS;           // [FUSE_MAX][cbd]
    double *frow;           // [ld]   running f row of the virtual table
    double *bvec;           // [n]    running b column of the virtual table
    int32_t *rowlab, *collab, *trace;
};

__global__ void __launch_bounds__(PRICE_THREADS, 1)
block_price_kernel(PriceArgs a) {
    __shared__ Scratch s;
    __shared__ LevelDiv s_lvl[FUSE_MAX];
    __shared__ double s_scal[FUSE_MAX];          // per level: COL_l[t] (row scans) or ROW_l[c] (column build)
    const int n = a.n, m = a.m, tid = threadIdx.x, nt = blockDim.x;
    const int64_t ld = a.ld, cbd = a.cbd;

    if (a.st->status != SPX_PIVOT) {             // sticky: nothing left to do in this pass
        if (tid == 0) a.plan->f = 0;
        return;
    }
    const int cur = (int)a.st->reserved[0] & 1;
    const double *A = a.A[cur];
    const int64_t npiv0 = a.st->npiv, cap = a.st->max_pivots;

    // running copies of the b column and the f row of the virtual table
    for (int t = tid; t < n; t += nt) a.bvec[t] = a.b[cur][t];
    for (int j = tid; j < m; j += nt) a.frow[j] = A[(int64_t)n * ld + j];
    __syncthreads();

    int status = SPX_PIVOT, f = 0, last_r = -1, last_c = -1, phase1 = 0;
    double last_p = 0.0;
    for (int i = 0; i < a.F; ++i) {
        // ---- K1: phase-1 row (:72-76) from the running b
        const int rb = block_first_index_fn(n, [&](int t) { return a.bvec[t]; }, IsNeg(), s);
        const int r1 = (rb == SPX_NONE) ? -1 : rb;
        int c;
        if (r1 >= 0) {
            // the row r1 of the virtual table: the stored row replayed through the pending levels (:82-85)
            if (tid < i) s_scal[tid] = a.COLS[(int64_t)tid * cbd + r1];
            __syncthreads();
            const double *row = A + (int64_t)r1 * ld;
            c = block_first_index_fn(m, [&](int j) {
                    double v = row[j];
                    for (int l = 0; l < i; ++l) v = apply_level(v, r1, j, s_lvl[l], a.ROWS[(int64_t)l * ld + j], s_scal[l]);
                    return v; }, IsPos(), s);
        } else if (a.rule == SPX_RULE_REFERENCE) {
            c = block_first_index_fn(m, [&](int j) { return a.frow[j]; }, IsNeg(), s);                 // :94-98
        } else {                                     // Dantzig: most negative, lowest index on ties
            unsigned long long best = ~0ull;
            for (int j = tid; j < m; j += nt) {
                const double v = a.frow[j];
                if (v < 0.0) { const unsigned long long k = orderable(v); best = k < best ? k : best; }
            }
            best = block_min_u64(best, s);
            int loc = SPX_NONE;
            if (best != ~0ull)
                for (int j = tid; j < m; j += nt) {
                    const double v = a.frow[j];
                    if (v < 0.0 && orderable(v) == best) { loc = j; break; }
                }
            c = block_min_int(loc, s);
        }
        if (c == SPX_NONE) { status = (r1 >= 0) ? SPX_INCORRECT : SPX_OPTIMAL; phase1 = (r1 >= 0); break; }

        // ---- the entering column of the virtual table: strided gather + replay; ratio test (:107-136)
        if (tid < i) s_scal[tid] = a.ROWS[(int64_t)tid * ld + c];
        __syncthreads();
        double *COLi = a.COLS + (int64_t)i * cbd;
        Ratio q = ratio_identity();
        const double *colp = A + c;
        for (int base = 0; base <= n; base += nt * PRICE_BATCH) {
            double v[PRICE_BATCH];
#pragma unroll
            for (int u = 0; u < PRICE_BATCH; ++u) {
                const int t = base + u * nt + tid;
                v[u] = (t <= n) ? colp[(int64_t)t * ld] : 0.0;
            }
#pragma unroll
            for (int u = 0; u < PRICE_BATCH; ++u) {
                const int t = base + u * nt + tid;
                if (t <= n) {
                    double w = v[u];
                    for (int l = 0; l < i; ++l) w = apply_level(w, t, c, s_lvl[l], s_scal[l], a.COLS[(int64_t)l * cbd + t]);
                    COLi[t] = w;
                    if (r1 < 0 && t < n) ratio_accumulate(q, t, w, a.bvec[t]);
                }
            }
        }
        int r;
        if (r1 >= 0) {
            r = r1;                                                                                   // :91
            __syncthreads();
        } else {
            q = block_ratio_reduce(q, s);
            bool elig_nan = false;
            if (q.elig_row != SPX_NONE) {
                const double v = __ddiv_rn(a.bvec[q.elig_row], COLi[q.elig_row]);
                elig_nan = (v != v);
            }
            r = ratio_decide(q, elig_nan);                                                            // :138-141
            if (r < 0) { status = SPX_NOCONV; break; }
        }
        const double p = COLi[r];
        if (npiv0 + i >= cap) { status = SPX_CAP; last_r = r; last_c = c; last_p = p; break; }

        // ---- the pivot row of the virtual table (contiguous read + replay)
        if (tid < i) s_scal[tid] = a.COLS[(int64_t)tid * cbd + r];
        __syncthreads();
        double *ROWi = a.ROWS + (int64_t)i * ld;
        const double *rowp = A + (int64_t)r * ld;
        for (int j = tid; j < m; j += nt) {
            double v = rowp[j];
            for (int l = 0; l < i; ++l) v = apply_level(v, r, j, s_lvl[l], a.ROWS[(int64_t)l * ld + j], s_scal[l]);
            ROWi[j] = v;
        }
        for (int j = m + tid; j < ld; j += nt) ROWi[j] = 0.0;          // padding columns stay zero
        // ---- record the level, then advance the running b column and f row by it
        if (tid == 0) {
            s_lvl[i].r = r; s_lvl[i].c = c; s_lvl[i].d = pivot_div_prepare(p);
            a.plan->lvl[i].r = r; a.plan->lvl[i].c = c; a.plan->lvl[i].p = p; a.plan->owner[i] = 0;
            const int32_t tmp = a.rowlab[c]; a.rowlab[c] = a.collab[r]; a.collab[r] = tmp;            // :152
            if (a.trace) { a.trace[2 * (npiv0 + i)] = r; a.trace[2 * (npiv0 + i) + 1] = c; }
        }
        __syncthreads();
        const LevelDiv L = s_lvl[i];
        const double br = a.bvec[r], fc = COLi[n];
        __syncthreads();
        for (int t = tid; t < n; t += nt) {
            const double bt = a.bvec[t];
            a.bvec[t] = (t == r) ? pivot_div(-bt, L.d) : cell_update(bt, L.d, br, COLi[t]);
        }
        for (int j = tid; j < m; j += nt) {
            const double fj = a.frow[j];
            a.frow[j] = (j == c) ? pivot_div(fc, L.d) : cell_update(fj, L.d, ROWi[j], fc);
        }
        __syncthreads();
        f = i + 1; last_r = r; last_c = c; last_p = p; phase1 = (r1 >= 0);
    }

    // ---- publish: the b column after f levels, the plan header, the state
    if (f > 0) for (int t = tid; t < n; t += nt) a.b[cur ^ 1][t] = a.bvec[t];
    if (tid == 0) {
        a.plan->f = f;
        a.plan->src = cur;
        spx_state *st = a.st;
        st->status = status; st->r = last_r; st->c = last_c; st->p = last_p;
        st->npiv = npiv0 + f; st->phase1 = phase1; st->slot = 0;
        st->hint_tag[0] = st->hint_tag[1] = -1;
        st->reserved[0] = (f > 0) ? (cur ^ 1) : cur;
    }
}

// ---- the same pricing spread over the whole GPU (cooperative launch) ---------------------------
// block_price_kernel is strictly sequential work for ONE CTA (0.3 ms at F = 4, 0.8 ms at F = 8 on
// cfg4) while the rest of the GPU idles between two fused passes.  Here every O(n) / O(m) loop is
// dealt out over all CTAs (one element per thread at cfg4) and the two decisions of a level — the
// entering column (an index min) and the leaving row (the ratio fold) — go through global scratch and
// a grid barrier: two barriers per level, ~10 us per level.  Data written by other CTAs during the
// kernel is read with ld.global.cg (L1 is not coherent across SMs).
constexpr int STAMPS = 6;                       // shard_price_kernel: %globaltimer per level at its phase boundaries
struct CoopScratch {
    int   idx[FUSE_MAX + 1][4];                 // per level: [first negative b, entering column, phase-1 column, -]
    unsigned long long key[FUSE_MAX + 1];       // Dantzig: most negative running f value (orderable image)
    unsigned int       arrive;                  // shard_price_kernel: CTAs whose candidate-column stores are fenced
    unsigned int       pad;
    unsigned long long bcast;                   // ... and the epoch of the last exchange result CTA 0 published
    unsigned long long stamp[FUSE_MAX + 1][STAMPS];   // debug: ns at level start / after sync 1 / column stored / keys
                                                //        exchanged / ratio partials synced / level done (thread 0)
    // persistent pricing engine (shard_price_kernel with npass > 1 passes per launch): the device-side hand-shake
    // between the ONE pricing kernel of an enqueue call and the update kernels of its passes
    unsigned long long plan_ready;              // the last pass number whose plan / ROW planes / COL planes are published
    unsigned long long upd_done;                // the last pass number whose update kernel has finished (its last CTA)
    unsigned long long p_alive;                 // first pass number of the pricing kernel that is resident right now
    unsigned int       upd_ctr;                 // CTAs of the running update kernel that are done
    unsigned int       abort;                   // a device-side wait of this call timed out: every later wait gives up at once
    unsigned long long go;                      // CTA 0 -> grid: (pass number << 1) | (its table is ready: 1, gave up: 0)
    unsigned long long pass_stamp[4];           // debug: ns at the last pass's start / table ready / first level / published
};

struct CoopArgs {
    PriceArgs  a;
    double    *bv[2];          // running b column, ping-pong per level
    CoopScratch *cs;
    Ratio     *part;           // [FUSE_MAX][gridDim.x] ratio partials
};

__global__ void __launch_bounds__(COOP_THREADS, 1)
coop_price_kernel(CoopArgs ca) {
    cg::grid_group grid = cg::this_grid();
    const PriceArgs &a = ca.a;
    __shared__ Scratch s;
    __shared__ LevelDiv s_lvl[FUSE_MAX];
    __shared__ double s_scal[FUSE_MAX];
    const int n = a.n, m = a.m, tid = threadIdx.x;
    const int G = gridDim.x, gtid = blockIdx.x * blockDim.x + tid, gn = G * blockDim.x;
    const int64_t ld = a.ld, cbd = a.cbd;

    if (a.st->status != SPX_PIVOT) {             // uniform over the grid: nobody reaches a barrier
        if (gtid == 0) a.plan->f = 0;
        return;
    }
    const int cur = (int)a.st->reserved[0] & 1;
    const double *A = a.A[cur];
    const int64_t npiv0 = a.st->npiv, cap = a.st->max_pivots;

    if (blockIdx.x == 0)
        for (int k = tid; k < (FUSE_MAX + 1) * 4; k += blockDim.x) {
            ca.cs->idx[k / 4][k % 4] = SPX_NONE;
            if (k % 4 == 0) ca.cs->key[k / 4] = ~0ull;
        }
    grid.sync();

    int status = SPX_PIVOT, f = 0, last_r = -1, last_c = -1, phase1 = 0;
    double last_p = 0.0;
    for (int i = 0; i <= a.F; ++i) {
        // ---------------- phase A: bring the running b column / f row to the virtual table k+i (for
        // i > 0 this also builds ROW_{i-1}, the pivot row of the level just chosen) and fold the
        // first-negative searches of level i into the same sweep
        double *bout = ca.bv[i & 1];
        int bneg = SPX_NONE, fneg = SPX_NONE;
        unsigned long long fkey = ~0ull;
        if (i == 0) {
            for (int t = gtid; t < n; t += gn) { const double v = a.b[cur][t]; bout[t] = v; if (v < 0.0) bneg = min(bneg, t); }
            for (int j = gtid; j < m; j += gn) {
                const double v = A[(int64_t)n * ld + j];
                a.frow[j] = v;
                if (v < 0.0) { fneg = min(fneg, j); const unsigned long long k = orderable(v); fkey = k < fkey ? k : fkey; }
            }
        } else {
            const LevelDiv L = s_lvl[i - 1];
            const double *bin = ca.bv[(i - 1) & 1];
            const double *COLL = a.COLS + (int64_t)(i - 1) * cbd;
            const double br = __ldcg(bin + L.r), fc = __ldcg(COLL + n);
            for (int t = gtid; t < n; t += gn) {
                const double bt = __ldcg(bin + t);
                const double v = (t == L.r) ? pivot_div(-bt, L.d) : cell_update(bt, L.d, br, __ldcg(COLL + t));
                bout[t] = v;
                if (v < 0.0) bneg = min(bneg, t);
            }
            if (tid < i - 1) s_scal[tid] = __ldcg(a.COLS + (int64_t)tid * cbd + L.r);
            __syncthreads();
            double *ROWL = a.ROWS + (int64_t)(i - 1) * ld;
            const double *rowp = A + (int64_t)L.r * ld;
            for (int j = gtid; j < ld; j += gn) {
                if (j >= m) { ROWL[j] = 0.0; continue; }                // padding columns stay zero
                double rv = rowp[j];
                for (int l = 0; l < i - 1; ++l)
                    rv = apply_level(rv, L.r, j, s_lvl[l], __ldcg(a.ROWS + (int64_t)l * ld + j), s_scal[l]);
                ROWL[j] = rv;
                const double fj = a.frow[j];                             // j is owned by this thread all kernel long
                const double v = (j == L.c) ? pivot_div(fc, L.d) : cell_update(fj, L.d, rv, fc);
                a.frow[j] = v;
                if (v < 0.0) { fneg = min(fneg, j); const unsigned long long k = orderable(v); fkey = k < fkey ? k : fkey; }
            }
        }
        if (i == a.F) { f = a.F; break; }                                // every level of the pass is chosen
        bneg = block_min_int(bneg, s);
        fneg = block_min_int(fneg, s);
        if (tid == 0) {
            if (bneg != SPX_NONE) atomicMin(&ca.cs->idx[i][0], bneg);
            if (fneg != SPX_NONE) atomicMin(&ca.cs->idx[i][1], fneg);
        }
        if (a.rule == SPX_RULE_DANTZIG) {
            fkey = block_min_u64(fkey, s);
            if (tid == 0 && fkey != ~0ull) atomicMin(&ca.cs->key[i], fkey);
        }
        grid.sync();
        const int rb = __ldcg(&ca.cs->idx[i][0]);
        const int r1 = (rb == SPX_NONE) ? -1 : rb;
        int c = __ldcg(&ca.cs->idx[i][1]);
        if (r1 >= 0) {
            // phase-1: first positive cell of the virtual row r1 (:82-85)
            if (tid < i) s_scal[tid] = __ldcg(a.COLS + (int64_t)tid * cbd + r1);
            __syncthreads();
            const double *row = A + (int64_t)r1 * ld;
            int loc = SPX_NONE;
            for (int j = gtid; j < m; j += gn) {
                double v = row[j];
                for (int l = 0; l < i; ++l) v = apply_level(v, r1, j, s_lvl[l], __ldcg(a.ROWS + (int64_t)l * ld + j), s_scal[l]);
                if (v > 0.0) { loc = j; break; }
            }
            loc = block_min_int(loc, s);
            if (tid == 0 && loc != SPX_NONE) atomicMin(&ca.cs->idx[i][2], loc);
            grid.sync();
            c = __ldcg(&ca.cs->idx[i][2]);
        } else if (a.rule == SPX_RULE_DANTZIG && c != SPX_NONE) {
            // lowest index among the columns that attain the most negative value
            const unsigned long long best = __ldcg(&ca.cs->key[i]);
            int loc = SPX_NONE;
            for (int j = gtid; j < m; j += gn) {
                const double v = a.frow[j];
                if (v < 0.0 && orderable(v) == best) { loc = j; break; }
            }
            loc = block_min_int(loc, s);
            if (tid == 0 && loc != SPX_NONE) atomicMin(&ca.cs->idx[i][2], loc);
            grid.sync();
            c = __ldcg(&ca.cs->idx[i][2]);
        }
        if (c == SPX_NONE) { status = (r1 >= 0) ? SPX_INCORRECT : SPX_OPTIMAL; phase1 = (r1 >= 0); f = i; break; }

        // ---------------- phase B: the entering column of the virtual table + the ratio fold (:107-136)
        if (tid < i) s_scal[tid] = __ldcg(a.ROWS + (int64_t)tid * ld + c);
        __syncthreads();
        double *COLi = a.COLS + (int64_t)i * cbd;
        Ratio q = ratio_identity();
        for (int t = gtid; t <= n; t += gn) {
            double w = A[(int64_t)t * ld + c];
            for (int l = 0; l < i; ++l) w = apply_level(w, t, c, s_lvl[l], s_scal[l], __ldcg(a.COLS + (int64_t)l * cbd + t));
            COLi[t] = w;
            if (r1 < 0 && t < n) ratio_accumulate(q, t, w, bout[t]);     // bout[t] was written by this very thread
        }
        q = block_ratio_reduce(q, s);
        Ratio *part = ca.part + (int64_t)i * G;
        if (tid == 0) part[blockIdx.x] = q;
        grid.sync();
        int r;
        if (r1 >= 0) {
            r = r1;                                                      // :91
        } else {
            Ratio z = ratio_identity();
            for (int k = tid; k < G; k += blockDim.x) {
                Ratio y;
                y.neg_val = __ldcg(&part[k].neg_val); y.neg_row = __ldcg(&part[k].neg_row);
                y.zero_row = __ldcg(&part[k].zero_row); y.elig_row = __ldcg(&part[k].elig_row);
                z = ratio_merge(z, y);
            }
            z = block_ratio_reduce(z, s);
            bool elig_nan = false;
            if (z.elig_row != SPX_NONE) {
                const double v = __ddiv_rn(__ldcg(bout + z.elig_row), __ldcg(COLi + z.elig_row));
                elig_nan = (v != v);
            }
            r = ratio_decide(z, elig_nan);                               // :138-141
            if (r < 0) { status = SPX_NOCONV; f = i; break; }
        }
        const double p = __ldcg(COLi + r);
        if (npiv0 + i >= cap) { status = SPX_CAP; last_r = r; last_c = c; last_p = p; f = i; break; }
        __syncthreads();
        if (tid == 0) {
            s_lvl[i].r = r; s_lvl[i].c = c; s_lvl[i].d = pivot_div_prepare(p);
            if (blockIdx.x == 0) {
                a.plan->lvl[i].r = r; a.plan->lvl[i].c = c; a.plan->lvl[i].p = p; a.plan->owner[i] = 0;
                const int32_t tmp = a.rowlab[c]; a.rowlab[c] = a.collab[r]; a.collab[r] = tmp;        // :152
                if (a.trace) { a.trace[2 * (npiv0 + i)] = r; a.trace[2 * (npiv0 + i) + 1] = c; }
            }
        }
        __syncthreads();
        last_r = r; last_c = c; last_p = p; phase1 = (r1 >= 0);
    }

    // ---------------- publish: b after f levels (phase A of iteration f wrote it), plan, state
    if (f > 0) {
        grid.sync();
        const double *bfin = ca.bv[f & 1];
        for (int t = gtid; t < n; t += gn) a.b[cur ^ 1][t] = __ldcg(bfin + t);
    }
    if (gtid == 0) {
        a.plan->f = f;
        a.plan->src = cur;
        spx_state *st = a.st;
        st->status = status; st->r = last_r; st->c = last_c; st->p = last_p;
        st->npiv = npiv0 + f; st->phase1 = phase1; st->slot = 0;
        st->hint_tag[0] = st->hint_tag[1] = -1;
        st->reserved[0] = (f > 0) ? (cur ^ 1) : cur;
    }
}

// ---- the cooperative pricing kernel for a COLUMN-SHARDED tableau (one process per GPU) ----------
// Rank g owns the columns [col0, col0 + m) of the body; b, labels, state and the plan are replicated.
// Per level the ranks exchange, straight from inside this kernel over NVLink peer memory:
//   column  every rank builds ITS best local entering column (gather + replay) and its threads store the
//           n+1 cells into its plane in every rank's XBOX;
//   key     then its 16-byte key + a release flag; every rank acquire-polls its local flags, takes the
//           lexicographic minimum and reads the winner's plane.
// The ratio test, the b column and the level bookkeeping are computed redundantly (bit-identically)
// on every rank.  XBOX of one rank:  COLS[3][FUSE_MAX][R][cbd] | keys[2][FUSE_MAX+1][2][R][2] |
// kflag[2][FUSE_MAX+1][2][R]  — one plane per (level, source rank): every rank that has a candidate stores
// its candidate column SPECULATIVELY together with its key (one exchange per level instead of a key
// round followed by a column round); the planes are triple-buffered by pass number (see the kernel),
// keys and flags double-buffered by pass parity.
constexpr int XB_MAX_RANKS = 16;

struct XBoxLayout {
    int64_t cols_off, keys_off, kflag_off, bytes;
};
__host__ __device__ inline XBoxLayout xbox_layout(int64_t cbd, int R) {
    XBoxLayout L;
    L.cols_off = 0;
    L.keys_off = (3LL * FUSE_MAX * R * cbd * 8 + 127) / 128 * 128;
    L.kflag_off = L.keys_off + (2LL * (FUSE_MAX + 1) * 2 * R * 16 + 127) / 128 * 128;
    L.bytes = L.kflag_off + (2LL * (FUSE_MAX + 1) * 2 * R * 8 + 127) / 128 * 128;
    return L;
}

struct ShardArgs {
    CoopArgs ca;                 // a.m = local columns, a.ld = local leading dimension, a.ROWS local planes
    int rank, R;
    int64_t col0;
    unsigned long long seq;      // pass number (1, 2, ...): the value the flags of this pass carry
    unsigned char *xbox[XB_MAX_RANKS];   // every rank's XBOX mapped into this process ([rank] = local)
    // the passes of this launch: pass q (0 <= q < npass) has number seq + q, writes plans[(seq + q) & 1] /
    // rows[(seq + q) & 1] and prices min(depth, pivots left) levels.  with_prev: the pass before the first one
    // is still being applied by its update kernel (look-ahead); within a launch every later pass has one.
    // persistent (npass may be > 1): the kernel stays resident for the whole enqueue call — it waits for the
    // update of pass q - 2 on cs->upd_done and publishes pass q on cs->plan_ready (see fused_run).
    PlanHeader *plans[2];
    double *rows[2];
    int npass, depth, with_prev, persistent, wait_updates;
    int64_t pivots;
};

__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long gtimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// Lazy replay of up to 2 * FUSE_MAX pending levels on ONE cell with every level's global operand in flight at once:
// `op[l] = ptr[l][idx]` does not depend on the running value, but a rolled loop issues load l + 1 only after
// apply_level(l) has consumed load l — up to 15 serialised L2 round trips per cell on the critical path of a level.
// ROWOP: the loaded operand is the level's ROW value (the scalar its COL value), or the other way round.
template <bool ROWOP>
__device__ __forceinline__ double replay_levels(double v, int t, int j, int nl, const LevelDiv *lvl,
                                                const double *const *ptr, int64_t idx, const double *scal) {
    double op[2 * FUSE_MAX];
#pragma unroll
    for (int l = 0; l < 2 * FUSE_MAX; ++l) op[l] = (l < nl) ? __ldcg(ptr[l] + idx) : 0.0;
#pragma unroll
    for (int l = 0; l < 2 * FUSE_MAX; ++l)
        if (l < nl) v = ROWOP ? apply_level(v, t, j, lvl[l], op[l], scal[l]) : apply_level(v, t, j, lvl[l], scal[l], op[l]);
    return v;
}

__device__ __forceinline__ void st_release_gpu_u64(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_gpu_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned int ld_acquire_gpu_u32(const unsigned int *p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// spin until *flag >= seq; false after 20 s (a peer died)
__device__ __forceinline__ bool wait_seq(const unsigned long long *flag, unsigned long long seq) {
    const unsigned long long t0 = gtimer_ns();
    while (ld_acquire_sys(flag) < seq) {
        if (gtimer_ns() - t0 > 20000000000ull) return false;
        __nanosleep(64);
    }
    return true;
}

// CTA 0 only: publish this rank's key of (level, kind) to every rank, collect everyone's, pick the
// lexicographic minimum.  Returns through sel[0] = global column (or SPX_NONE), sel[1] = owner rank,
// sel[2] = 0 / 1 (timeout).
__device__ void exchange_keys(const ShardArgs &sa, const XBoxLayout &XL, int kpar, int level, int kind,
                              unsigned long long kh, unsigned long long kl, unsigned long long seq, int *sel) {
    const int tid = threadIdx.x, R = sa.R;
    level += kpar * (FUSE_MAX + 1);                     // the slot set of this pass parity
    __shared__ int s_to;
    if (tid == 0) s_to = 0;
    __syncthreads();
    if (tid < R) {
        unsigned char *box = sa.xbox[tid];
        unsigned long long *slot = reinterpret_cast<unsigned long long *>(box + XL.keys_off) +
                                   (((int64_t)level * 2 + kind) * R + sa.rank) * 2;
        slot[0] = kh; slot[1] = kl;
        unsigned long long *fl = reinterpret_cast<unsigned long long *>(box + XL.kflag_off) +
                                 ((int64_t)level * 2 + kind) * R + sa.rank;
        st_release_sys(fl, seq);                         // release: the two key words are visible first
        // ... and wait for rank `tid`'s key in MY box
        const unsigned long long *mine = reinterpret_cast<const unsigned long long *>(sa.xbox[sa.rank] + XL.kflag_off) +
                                         ((int64_t)level * 2 + kind) * R + tid;
        if (!wait_seq(mine, seq)) s_to = 1;
    }
    __syncthreads();
    if (tid == 0) {
        unsigned long long bh = ~0ull, bl = ~0ull; int win = -1;
        const unsigned long long *keys = reinterpret_cast<const unsigned long long *>(sa.xbox[sa.rank] + XL.keys_off) +
                                         ((int64_t)level * 2 + kind) * R * 2;
        for (int g = 0; g < R; ++g) {
            const unsigned long long h = __ldcg(keys + 2 * g), l = __ldcg(keys + 2 * g + 1);
            if (l != ~0ull && (win < 0 || h < bh || (h == bh && l < bl))) { bh = h; bl = l; win = g; }
        }
        sel[0] = (win < 0) ? SPX_NONE : (int)bl;
        sel[1] = win;
        sel[2] = s_to;
    }
}

// The device-side waits of the persistent pricing engine.  Every one of them gives up after ENGINE_WAIT_NS or as soon
// as another wait of the same call has given up (cs->abort): a lost hand-shake ends the call with SPX_PEER_TIMEOUT
// (and a table that is no longer meaningful) instead of hanging the GPU.
constexpr unsigned long long ENGINE_WAIT_NS = 30000000000ull;
__device__ __forceinline__ unsigned int ld_relaxed_gpu_u32(const unsigned int *p) {
    unsigned int v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// one thread: until *flag >= v (acquire, gpu scope: the writer is another kernel of THIS device); false = gave up
__device__ __forceinline__ bool spin_ge_or_abort(const unsigned long long *flag, unsigned long long v, unsigned int *abort) {
    const unsigned long long t0 = gtimer_ns();
    for (unsigned it = 0;; ++it) {
        if (ld_acquire_gpu_u64(flag) >= v) return true;
        if ((it & 15u) == 15u) {
            if (ld_relaxed_gpu_u32(abort) != 0u) return false;
            if (gtimer_ns() - t0 > ENGINE_WAIT_NS) { atomicExch(abort, 1u); return false; }
        }
        __nanosleep(40);
    }
}
// the whole CTA: true when *flag >= v was observed by thread 0
__device__ __forceinline__ bool cta_wait_ge(const unsigned long long *flag, unsigned long long v, unsigned int *abort) {
    int ok = 1;
    if (threadIdx.x == 0) ok = spin_ge_or_abort(flag, v, abort) ? 1 : 0;
    return __syncthreads_and(ok) != 0;
}
// pricing kernel, pass `seq`: CTA 0 waits for the update of pass seq - 2 and tells the grid how that went, so that
// every CTA takes the same decision (they meet at grid barriers right after)
__device__ __forceinline__ bool grid_gate_update_done(CoopScratch *cs, unsigned long long seq) {
    int ok = 1;
    if (threadIdx.x == 0) {
        if (blockIdx.x == 0) {
            ok = spin_ge_or_abort(&cs->upd_done, seq - 2ull, &cs->abort) ? 1 : 0;
            __threadfence();
            st_release_gpu_u64(&cs->go, (seq << 1) | (unsigned long long)ok);
        } else {
            unsigned long long g;
            while (((g = ld_acquire_gpu_u64(&cs->go)) >> 1) != seq) __nanosleep(40);
            ok = (int)(g & 1ull);
        }
    }
    return __syncthreads_and(ok) != 0;
}

__global__ void __launch_bounds__(COOP_THREADS, 1)
shard_price_kernel(ShardArgs sa) {
    cg::grid_group grid = cg::this_grid();
    const CoopArgs &ca = sa.ca;
    const PriceArgs &a = ca.a;
    __shared__ Scratch s;
    // look-ahead: levels 0..np-1 are the PREVIOUS pass's (still being applied to the stored table by its
    // update kernel while this kernel runs), levels np..np+i-1 are this pass's
    __shared__ LevelDiv s_lvl[2 * FUSE_MAX];
    __shared__ double s_scal[2 * FUSE_MAX];
    __shared__ const double *s_rowp[2 * FUSE_MAX];       // per level: its ROW plane (local columns)
    __shared__ const double *s_colp[2 * FUSE_MAX];       // per level: the winner's COL plane in MY box
    const int n = a.n, m = a.m, tid = threadIdx.x;
    const int G = gridDim.x, gtid = blockIdx.x * blockDim.x + tid, gn = G * blockDim.x;
    const int64_t ld = a.ld, cbd = a.cbd, col0 = sa.col0;
    const XBoxLayout XL = xbox_layout(cbd, sa.R);
    int *gsel = &ca.cs->idx[FUSE_MAX][0];               // CTA 0 -> grid: {column, owner, timeout} of the last exchange

    // the solve's state travels in registers from pass to pass (every thread derives the same values); the device
    // copy is written once per pass by thread 0 for the host and for the next launch
    int status = (int)a.st->status;
    int cur = (int)a.st->reserved[0] & 1;                // the buffer the next pass's INPUT table lives in
    int64_t npiv = a.st->npiv;
    const int64_t cap = a.st->max_pivots;
    int64_t left = sa.pivots;
    int prev_f = 0;                                      // levels of the previous pass of THIS launch

    if (sa.persistent) {
        // every CTA of this kernel is resident now: the update kernels of this call may start (they spin on
        // plan_ready while they occupy the other SMs — had they started first, this kernel could not be placed)
        grid.sync();
        if (gtid == 0) st_release_gpu_u64(&ca.cs->p_alive, sa.seq);
    }

    for (int q = 0; q < sa.npass; ++q) {
    const unsigned long long seq = sa.seq + (unsigned long long)q;
    const int hq = (int)(seq & 1ull);
    PlanHeader *plan = sa.plans[hq];
    double *ROWS = sa.rows[hq];
    const bool has_prev = (q > 0) || sa.with_prev;
    const PlanHeader *prev_plan = has_prev ? sa.plans[hq ^ 1] : nullptr;
    const double *prev_ROWS = has_prev ? sa.rows[hq ^ 1] : nullptr;
    const int F = (int)(left < (int64_t)sa.depth ? left : (int64_t)sa.depth);
    left -= F;
    const bool last_pass = (q == sa.npass - 1);
    if (last_pass && gtid == 0) ca.cs->pass_stamp[0] = gtimer_ns();
    // the update of pass q - 2 read plan[hq], ROWS[hq] and wrote the table this pass gathers from
    if (sa.persistent && sa.wait_updates && q >= 2 && !grid_gate_update_done(ca.cs, seq) && status == SPX_PIVOT)
        status = SPX_PEER_TIMEOUT;
    if (last_pass && gtid == 0) ca.cs->pass_stamp[1] = gtimer_ns();

    // COL planes are TRIPLE-buffered by pass number: with look-ahead a fast rank can be pricing pass q+1
    // (storing into every rank's planes) while a slow rank's update of pass q-1 still reads its planes —
    // the slow rank's pricing of pass q, which the fast rank had to wait for, runs concurrently with that
    // update.  Pass q+2 cannot start anywhere before every rank's update q-1 has finished.
    const int par = (int)(seq % 3ull);
    const int par_prev = (int)((seq + 2ull) % 3ull);
    const int kpar = (int)(seq & 1ull);                  // key / flag slots: double-buffered by pass parity
    // this pass's COLS planes in MY box: plane (level l, source rank g) at COLS + (l * R + g) * cbd
    double *COLS = reinterpret_cast<double *>(sa.xbox[sa.rank] + XL.cols_off) + (int64_t)par * FUSE_MAX * sa.R * cbd;

    if (status != SPX_PIVOT) {                           // uniform over the grid AND over the ranks
        if (gtid == 0) {
            plan->f = 0;
            a.st->status = status;                       // (a wait that gave up above ends the solve here)
            if (sa.persistent) { __threadfence(); st_release_gpu_u64(&ca.cs->plan_ready, seq); }
        }
        prev_f = 0;
        continue;
    }
    // cur = the buffer this pass's INPUT table lives in (it may still be under construction by the
    // previous pass's update); the cells are gathered from the last MATERIALISED table: the other buffer
    // when np levels of the previous pass are pending, cur itself otherwise
    // np: within a launch the previous pass's f is known to every thread (its plan->f store is not ordered
    // before this pass by any grid barrier); across launches it comes from memory
    const int np = (prev_plan == nullptr) ? 0 : (q > 0 ? prev_f : __ldcg(&prev_plan->f));
    const double *A = a.A[np > 0 ? (cur ^ 1) : cur];
    const int64_t npiv0 = npiv;
    __syncthreads();                                     // the previous pass's readers of s_lvl / s_colp / s_rowp are done
    if (tid < np) {
        const int64_t cl = (int64_t)__ldcg(&prev_plan->lvl[tid].c) - col0;
        s_lvl[tid].r = __ldcg(&prev_plan->lvl[tid].r);
        s_lvl[tid].c = (cl >= 0 && cl < m) ? (int)cl : -1;
        s_lvl[tid].d = pivot_div_prepare(__ldcg(&prev_plan->lvl[tid].p));
        s_rowp[tid] = prev_ROWS + (int64_t)tid * ld;
        s_colp[tid] = reinterpret_cast<const double *>(sa.xbox[sa.rank] + XL.cols_off) +
                      (((int64_t)par_prev * FUSE_MAX + tid) * sa.R + __ldcg(&prev_plan->owner[tid])) * cbd;
    }
    __syncthreads();

    if (blockIdx.x == 0)
        for (int k = tid; k < (FUSE_MAX + 1) * 4; k += blockDim.x) {
            ca.cs->idx[k / 4][k % 4] = SPX_NONE;
            if (k % 4 == 0) ca.cs->key[k / 4] = ~0ull;
        }
    grid.sync();
    if (last_pass && gtid == 0) ca.cs->pass_stamp[2] = gtimer_ns();

    int f = 0, last_r = -1, last_c = -1, phase1 = 0;
    double last_p = 0.0;
    for (int i = 0; i <= F; ++i) {
        // ---------------- phase A (see coop_price_kernel): running b (replicated), local f row shard,
        // local shard of ROW_{i-1}, first-negative folds
        double *bout = ca.bv[i & 1];
        int bneg = SPX_NONE, fneg = SPX_NONE;
        unsigned long long fkey = ~0ull;
        if (i == 0) {
            for (int t = gtid; t < n; t += gn) { const double v = __ldcg(&a.b[cur][t]); bout[t] = v; if (v < 0.0) bneg = min(bneg, t); }
            for (int j = gtid; j < m; j += gn) {
                // np > 0: the previous pricing left the running f row at exactly this table
                const double v = (np > 0) ? __ldcg(&a.frow[j]) : __ldcg(&A[(int64_t)n * ld + j]);
                a.frow[j] = v;
                if (v < 0.0) { fneg = min(fneg, j); const unsigned long long k = orderable(v); fkey = k < fkey ? k : fkey; }
            }
        } else {
            const LevelDiv L = s_lvl[np + i - 1];        // L.c is the LOCAL index of the pivot column or -1
            const double *bin = ca.bv[(i - 1) & 1];
            const double *COLL = s_colp[np + i - 1];
            const double br = __ldcg(bin + L.r), fc = __ldcg(COLL + n);
            for (int t = gtid; t < n; t += gn) {
                const double bt = __ldcg(bin + t);
                const double v = (t == L.r) ? pivot_div(-bt, L.d) : cell_update(bt, L.d, br, __ldcg(COLL + t));
                bout[t] = v;
                if (v < 0.0) bneg = min(bneg, t);
            }
            if (tid < np + i - 1) s_scal[tid] = __ldcg(s_colp[tid] + L.r);
            __syncthreads();
            double *ROWL = ROWS + (int64_t)(i - 1) * ld;
            const double *rowp = A + (int64_t)L.r * ld;
            for (int j = gtid; j < ld; j += gn) {
                if (j >= m) { ROWL[j] = 0.0; continue; }
                const double rv = replay_levels<true>(__ldcg(rowp + j), L.r, j, np + i - 1, s_lvl, s_rowp, j, s_scal);
                ROWL[j] = rv;
                const double fj = __ldcg(&a.frow[j]);
                const double v = (j == L.c) ? pivot_div(fc, L.d) : cell_update(fj, L.d, rv, fc);
                a.frow[j] = v;
                if (v < 0.0) { fneg = min(fneg, j); const unsigned long long k = orderable(v); fkey = k < fkey ? k : fkey; }
            }
        }
        if (i == F) { f = F; break; }
        if (gtid == 0) ca.cs->stamp[i][0] = gtimer_ns();
        bneg = block_min_int(bneg, s);
        fneg = block_min_int(fneg, s);
        if (tid == 0) {
            if (bneg != SPX_NONE) atomicMin(&ca.cs->idx[i][0], bneg);
            if (fneg != SPX_NONE) atomicMin(&ca.cs->idx[i][1], fneg);
        }
        if (a.rule == SPX_RULE_DANTZIG) {
            fkey = block_min_u64(fkey, s);
            if (tid == 0 && fkey != ~0ull) atomicMin(&ca.cs->key[i], fkey);
        }
        grid.sync();
        if (gtid == 0) ca.cs->stamp[i][1] = gtimer_ns();
        const int rb = __ldcg(&ca.cs->idx[i][0]);
        const int r1 = (rb == SPX_NONE) ? -1 : rb;       // identical on every rank: b is replicated
        int cloc = __ldcg(&ca.cs->idx[i][1]);            // this rank's candidate (local index)
        unsigned long long kh = 0ull;
        int kind = 0;
        if (r1 >= 0) {
            // phase-1: first positive cell of the local part of the virtual row r1 (:82-85)
            kind = 1;
            if (tid < np + i) s_scal[tid] = __ldcg(s_colp[tid] + r1);
            __syncthreads();
            const double *row = A + (int64_t)r1 * ld;
            int loc = SPX_NONE;
            for (int j = gtid; j < m; j += gn) {
                const double v = replay_levels<true>(__ldcg(row + j), r1, j, np + i, s_lvl, s_rowp, j, s_scal);
                if (v > 0.0) { loc = j; break; }
            }
            loc = block_min_int(loc, s);
            if (tid == 0 && loc != SPX_NONE) atomicMin(&ca.cs->idx[i][2], loc);
            grid.sync();
            cloc = __ldcg(&ca.cs->idx[i][2]);
        } else if (a.rule == SPX_RULE_DANTZIG && cloc != SPX_NONE) {
            const unsigned long long best = __ldcg(&ca.cs->key[i]);
            int loc = SPX_NONE;
            for (int j = gtid; j < m; j += gn) {
                const double v = __ldcg(&a.frow[j]);
                if (v < 0.0 && orderable(v) == best) { loc = j; break; }
            }
            loc = block_min_int(loc, s);
            if (tid == 0 && loc != SPX_NONE) atomicMin(&ca.cs->idx[i][2], loc);
            grid.sync();
            cloc = __ldcg(&ca.cs->idx[i][2]);
            kh = best;
        }
        // ---------------- phase B: this rank's candidate column of the virtual table, stored straight
        // into its plane in EVERY rank's XBOX (speculative: only the winner's plane will be read)
        if (cloc != SPX_NONE) {
            if (tid < np + i) s_scal[tid] = __ldcg(s_rowp[tid] + cloc);
            __syncthreads();
            const int64_t plane = (((int64_t)par * FUSE_MAX + i) * sa.R + sa.rank) * cbd;
            for (int t = gtid; t <= n; t += gn) {
                const double w = replay_levels<false>(__ldcg(&A[(int64_t)t * ld + cloc]), t, cloc, np + i, s_lvl, s_colp, t, s_scal);
                for (int g = 0; g < sa.R; ++g)
                    (reinterpret_cast<double *>(sa.xbox[g] + XL.cols_off) + plane)[t] = w;
            }
        }
        // ---------------- one exchange: key + flag.  Instead of two grid barriers around it: every CTA fences its
        // column stores and ARRIVES at a counter; CTA 0 waits for all of them (the stores are then ordered before
        // the flag it releases to the peers), exchanges, and BROADCASTS the result under a per-level epoch the
        // other CTAs spin on — an all-to-one plus a one-to-all instead of two all-to-alls.
        {
            const unsigned long long epoch = seq * (unsigned long long)(2 * FUSE_MAX) + (unsigned long long)i + 1ull;
            __threadfence_system();
            __syncthreads();
            if (blockIdx.x == 0) {
                if (tid == 0) {
                    atomicAdd(&ca.cs->arrive, 1u);
                    while (ld_acquire_gpu_u32(&ca.cs->arrive) < (unsigned)G) { }
                    ca.cs->arrive = 0u;                      // nobody arrives again before the broadcast below
                    ca.cs->stamp[i][2] = gtimer_ns();
                }
                __syncthreads();
                exchange_keys(sa, XL, kpar, i, kind, (cloc == SPX_NONE) ? ~0ull : kh,
                              (cloc == SPX_NONE) ? ~0ull : (unsigned long long)(col0 + cloc), seq, gsel);
                __syncthreads();
                if (tid == 0) { __threadfence(); st_release_gpu_u64(&ca.cs->bcast, epoch); }
            } else if (tid == 0) {
                __threadfence();
                atomicAdd(&ca.cs->arrive, 1u);
                while (ld_acquire_gpu_u64(&ca.cs->bcast) != epoch) { }
            }
            __syncthreads();
        }
        if (gtid == 0) ca.cs->stamp[i][3] = gtimer_ns();
        const int c = __ldcg(gsel + 0), owner = __ldcg(gsel + 1);
        if (__ldcg(gsel + 2)) { status = SPX_PEER_TIMEOUT; f = i; break; }
        if (c == SPX_NONE) { status = (r1 >= 0) ? SPX_INCORRECT : SPX_OPTIMAL; phase1 = (r1 >= 0); f = i; break; }
        const int64_t cl64 = (int64_t)c - col0;
        const int clocal = (cl64 >= 0 && cl64 < m) ? (int)cl64 : -1;
        if (tid == 0) {
            s_colp[np + i] = COLS + ((int64_t)i * sa.R + owner) * cbd;
            s_rowp[np + i] = ROWS + (int64_t)i * ld;
        }
        __syncthreads();

        // ---------------- ratio fold on the received column (every rank, identical) (:107-136)
        const double *COLi = s_colp[np + i];
        Ratio qr = ratio_identity();
        if (r1 < 0)
            for (int t = gtid; t < n; t += gn) ratio_accumulate(qr, t, __ldcg(COLi + t), bout[t]);
        qr = block_ratio_reduce(qr, s);
        Ratio *part = ca.part + (int64_t)i * G;
        if (tid == 0) part[blockIdx.x] = qr;
        grid.sync();
        if (gtid == 0) ca.cs->stamp[i][4] = gtimer_ns();
        int r;
        if (r1 >= 0) {
            r = r1;
        } else {
            Ratio z = ratio_identity();
            for (int k = tid; k < G; k += blockDim.x) {
                Ratio y;
                y.neg_val = __ldcg(&part[k].neg_val); y.neg_row = __ldcg(&part[k].neg_row);
                y.zero_row = __ldcg(&part[k].zero_row); y.elig_row = __ldcg(&part[k].elig_row);
                z = ratio_merge(z, y);
            }
            z = block_ratio_reduce(z, s);
            bool elig_nan = false;
            if (z.elig_row != SPX_NONE) {
                const double v = __ddiv_rn(__ldcg(bout + z.elig_row), __ldcg(COLi + z.elig_row));
                elig_nan = (v != v);
            }
            r = ratio_decide(z, elig_nan);
            if (r < 0) { status = SPX_NOCONV; f = i; break; }
        }
        const double p = __ldcg(COLi + r);
        if (npiv0 + i >= cap) { status = SPX_CAP; last_r = r; last_c = c; last_p = p; f = i; break; }
        __syncthreads();
        if (tid == 0) {
            s_lvl[np + i].r = r; s_lvl[np + i].c = clocal; s_lvl[np + i].d = pivot_div_prepare(p);
            if (blockIdx.x == 0) {
                ca.cs->stamp[i][5] = gtimer_ns();
                plan->lvl[i].r = r; plan->lvl[i].c = c; plan->lvl[i].p = p;             // GLOBAL column in the plan
                plan->owner[i] = owner;
                const int32_t tmp = a.rowlab[c]; a.rowlab[c] = a.collab[r]; a.collab[r] = tmp;
                if (a.trace) { a.trace[2 * (npiv0 + i)] = r; a.trace[2 * (npiv0 + i) + 1] = c; }
            }
        }
        __syncthreads();
        last_r = r; last_c = c; last_p = p; phase1 = (r1 >= 0);
    }

    // ---------------- publish the pass: b after f levels, plan, state — then (persistent) release the update
    if (f > 0) {
        grid.sync();                                     // every CTA's ROW planes and running b are complete
        const double *bfin = ca.bv[f & 1];
        for (int t = gtid; t < n; t += gn) a.b[cur ^ 1][t] = __ldcg(bfin + t);
    }
    if (gtid == 0) {
        plan->f = f;
        plan->src = cur;
        spx_state *st = a.st;
        st->status = status; st->r = last_r; st->c = last_c; st->p = last_p;
        st->npiv = npiv0 + f; st->phase1 = phase1; st->slot = 0;
        st->hint_tag[0] = st->hint_tag[1] = -1;
        st->reserved[0] = (f > 0) ? (cur ^ 1) : cur;
        if (sa.persistent) { __threadfence(); st_release_gpu_u64(&ca.cs->plan_ready, seq); }
        if (last_pass) ca.cs->pass_stamp[3] = gtimer_ns();
    }
    npiv = npiv0 + f;
    prev_f = f;
    if (f > 0) cur ^= 1;
    }   // passes
}

// ---- the fused streaming update: one pass over the body applies plan->f levels -----------------
struct FusedSmem {
    double rows[FUSE_MAX][FUP_TC];      // per level: slice of ROW_l for this column tile
    double cols[FUSE_MAX][FUP_TR];      // per level: slice of COL_l for this row tile
};

template <int MINB>
__global__ void __launch_bounds__(FUP_THREADS, MINB)
update_fused_kernel(double *A0, double *A1, int n, int m, int64_t ld, int64_t cbd, int64_t col0, int R,
                    const PlanHeader *__restrict__ plan, const double *__restrict__ ROWS,
                    const double *__restrict__ COLS) {
    const int f = plan->f;
    if (f <= 0) return;
    extern __shared__ __align__(128) unsigned char fus_raw[];
    FusedSmem &sm = *reinterpret_cast<FusedSmem *>(fus_raw);
    __shared__ LevelDiv s_lvl[FUSE_MAX];
    __shared__ alignas(8) uint64_t s_bar;

    const int src = plan->src;
    const double *__restrict__ Ain = src ? A1 : A0;
    double *__restrict__ Aout = src ? A0 : A1;

    const int tid = threadIdx.x;
    const int j0 = blockIdx.x * FUP_TC;
    const int i0 = blockIdx.y * FUP_TR;
    const int rows = min(FUP_TR, n + 1 - i0);
    const uint32_t row_bytes = (uint32_t)(min((int64_t)FUP_TC, ld - j0) * 8);
    const uint32_t col_bytes = (uint32_t)(FUP_TR * 8);                 // COLS planes are padded to whole tiles
    if (tid == 0) {
        mbar_init(&s_bar, 1);
        mbar_fence_init();
        mbar_expect_tx(&s_bar, (uint32_t)f * (row_bytes + col_bytes));
        for (int l = 0; l < f; ++l) {
            bulk_g2s(sm.rows[l], ROWS + (int64_t)l * ld + j0, row_bytes, &s_bar);
            bulk_g2s(sm.cols[l], COLS + ((int64_t)l * R + plan->owner[l]) * cbd + i0, col_bytes, &s_bar);
        }
    }
    if (tid < f) {
        s_lvl[tid].r = plan->lvl[tid].r;
        // the plan carries GLOBAL column indices; this shard owns [col0, col0 + m)
        const int64_t cl = (int64_t)plan->lvl[tid].c - col0;
        s_lvl[tid].c = (cl >= 0 && cl < m) ? (int)cl : -1;
        s_lvl[tid].d = pivot_div_prepare(plan->lvl[tid].p);
    }
    __syncthreads();
    mbar_wait(&s_bar, 0);

    const int j = j0 + 2 * tid;
    if (j >= m) return;
    // CTA-uniform: does any level's pivot row / column cross this tile?
    bool special = false;
    for (int l = 0; l < f; ++l) {
        const int rl = s_lvl[l].r, cl = s_lvl[l].c;
        special = special || (rl >= i0 && rl < i0 + rows) || (cl >= j0 && cl < j0 + FUP_TC);
    }
    const double *srcp = Ain + (int64_t)i0 * ld + j;
    double *dstp = Aout + (int64_t)i0 * ld + j;
    for (int ii = 0; ii < rows; ii += FUP_UNROLL) {
        double2 t[FUP_UNROLL];
#pragma unroll
        for (int u = 0; u < FUP_UNROLL; ++u)
            if (ii + u < rows) t[u] = ld_stream(srcp + (int64_t)(ii + u) * ld);
        if (!special) {
            // steady state: F dependent rank-1 updates per cell, guards accumulated, no branches
            bool ok = true;
            for (int l = 0; l < f; ++l) {
                const double2 rj = *reinterpret_cast<const double2 *>(&sm.rows[l][2 * tid]);
                const PivotDiv d = s_lvl[l].d;
#pragma unroll
                for (int u = 0; u < FUP_UNROLL; ++u) {
                    const double ci = sm.cols[l][ii + u];
                    t[u].x = cell_update_unchecked(t[u].x, d, rj.x, ci, ok);
                    t[u].y = cell_update_unchecked(t[u].y, d, rj.y, ci, ok);
                }
            }
            if (__builtin_expect(!ok, 0)) {
                // some quotient left the fast path's exponent range (an exact zero, a denormal...):
                // redo this thread's batch from the stored cells with the fully guarded division
#pragma unroll
                for (int u = 0; u < FUP_UNROLL; ++u)
                    if (ii + u < rows) t[u] = ld_stream(srcp + (int64_t)(ii + u) * ld);
                for (int l = 0; l < f; ++l) {
                    const double2 rj = *reinterpret_cast<const double2 *>(&sm.rows[l][2 * tid]);
                    const PivotDiv d = s_lvl[l].d;
#pragma unroll
                    for (int u = 0; u < FUP_UNROLL; ++u) {
                        const double ci = sm.cols[l][ii + u];
                        t[u].x = cell_update(t[u].x, d, rj.x, ci);
                        t[u].y = cell_update(t[u].y, d, rj.y, ci);
                    }
                }
            }
        } else {
            for (int l = 0; l < f; ++l) {
                const double2 rj = *reinterpret_cast<const double2 *>(&sm.rows[l][2 * tid]);
                const LevelDiv L = s_lvl[l];
#pragma unroll
                for (int u = 0; u < FUP_UNROLL; ++u) {
                    const int ti = i0 + ii + u;
                    const double ci = sm.cols[l][ii + u];
                    t[u].x = apply_level(t[u].x, ti, j, L, rj.x, ci);
                    t[u].y = apply_level(t[u].y, ti, j + 1, L, rj.y, ci);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < FUP_UNROLL; ++u)
            if (ii + u < rows) st_stream(dstp + (int64_t)(ii + u) * ld, t[u]);
    }
}

// ---- K6b: the fused update with the LAZY range guard (the default; SPX_OPT_FUSE_VARIANT = 1 selects the
// round-1 kernel above).  Same arithmetic, same bits; what changed is what the steady state pays per cell-level.
//
// Round 1's kernel spent 3 integer instructions per quotient on the accumulated range test of
// pivot_div_unchecked.  Measured on a B200 with everything in registers (tools/fp64_lab.cu,
// profiles/r2/r2c_fp64_lab.log): the bare 6-instruction fp64 chain holds the fp64 pipe at 92 %, the chain plus
// that guard at 78 %, plus a one-instruction-per-two-cells FMNMX3 guard at 82 % — the guard, not latency, was the
// gap (the register-prefetch / tall-tile schedules of round 1's experimental kernel measured no faster,
// profiles/r2/r2b_fused_update_variant_sweep.log, and are gone).  Here the per-level guard is GONE:
//
//   the fast division (q0 = a*y; rem = fma(-p, q0, a); q = fma(y, rem, q0), y the correctly rounded reciprocal)
//   returns RN(a / p) unless (i) a, q0 or q overflow — then q is Inf/NaN, and Inf/NaN stay Inf/NaN through every
//   later level — or (ii) |a| < 2^-969 or q is subnormal (rem or q underflow, or a = +-0 whose quotient gets the
//   wrong SIGN of zero).  In case (ii) the computed q~ and the true q are both tiny: |q|, |q~| <= 2^-969 / |p|.
//   Let every pivot of the pass satisfy 2^-100 <= |p| <= 2^100 (checked per tile; otherwise the guarded path
//   runs).  A pair (q~, q) with |q|, |q~| <= B entering the next level gives a = RN(RN(t p') - rc): either
//   |rc| >= 2^54 B |p'|, then t p' is absorbed entirely, a~ = a = -rc and the error is gone; or |rc| is smaller,
//   then |a|, |a~| < 2^55 B |p'| and the pair stays small with B' = 2^55 B.  Starting from B = 2^-869, seven more
//   levels give B <= 2^-484.  So a value that differs from the reference's is, at the END of the pass, either
//   Inf/NaN or smaller than 2^-484 in magnitude: ONE range test on the outputs, 2^-400 <= |q| < 2^1009, proves
//   all FUSE_MAX levels of a cell exact.  A thread whose batch fails it re-does the batch from the stored cells
//   with the fully guarded division (exact zeros — sparse tableaus — take that road, as they did in round 1).
//   spx_selftest_lazy_guard() runs adversarial chains (zeros, subnormals, cancellation to 2^-1000, Inf) through
//   both paths on the device and compares bits; tests/test_gpu_parity.py calls it.
//
// The test itself is two FMNMX3 per two outputs on the sign-stripped HIGH WORDS read as fp32 (monotonic in the
// fp64 magnitude; max.NaN keeps the NaN pattern of exponents >= 2040): 16 instructions per 768 fp64 issues.
// Other changes: the level loop is fully unrolled for a full pass (f == FUSE_MAX); (p, y) of a level is one
// 128-bit shared load, the 8 column multipliers four; the slow path (pivot row / pivot column inside the batch)
// is taken per 8-row batch and per warp instead of per tile; the first batch's loads are issued before the CTA
// waits for its staged slices.
constexpr unsigned LZ_LOW  = (unsigned)(1023 - 400) << 20;     // outputs below 2^-400 are re-done exactly
constexpr unsigned LZ_HIGH = 0x7f000000u;                      // ... and so are |q| >= 2^1009, Inf, NaN
constexpr int      LZ_P_SPAN = 100;                            // lazy guard only if 2^-100 <= |p| <= 2^100

__device__ __forceinline__ double cell_update_raw(double t, double p, double y, double rj, double ci) {
    const double a   = __dsub_rn(__dmul_rn(t, p), __dmul_rn(rj, ci));      // :173-175, three roundings
    const double q0  = __dmul_rn(a, y);
    const double rem = __fma_rn(-p, q0, a);
    return __fma_rn(y, rem, q0);
}

__device__ __forceinline__ void range_fold(float &lo, float &hi, double x, double y) {
    const float fx = fabsf(__int_as_float(__double2hiint(x))), fy = fabsf(__int_as_float(__double2hiint(y)));
    asm("min.f32 %0, %0, %1, %2;" : "+f"(lo) : "f"(fx), "f"(fy));                  // FMNMX3
    asm("max.NaN.f32 %0, %0, %1, %2;" : "+f"(hi) : "f"(fx), "f"(fy));              // FMNMX3.NAN
}
__device__ __forceinline__ bool range_ok(float lo, float hi) {
    return __float_as_uint(lo) >= LZ_LOW && __float_as_uint(hi) <= LZ_HIGH;
}

__device__ __forceinline__ void load_cols8(double (&cv)[FUP_UNROLL], const double *p) {
    static_assert(FUP_UNROLL == 8, "four 128-bit shared loads");
    const double2 *p2 = reinterpret_cast<const double2 *>(p);
#pragma unroll
    for (int u = 0; u < 4; ++u) { const double2 v = p2[u]; cv[2 * u] = v.x; cv[2 * u + 1] = v.y; }
}

// one level on a thread's 8 x 2 cells, no guard
__device__ __forceinline__ void lazy_level(double2 (&t)[FUP_UNROLL], const double2 rj, const double2 py, const double *cols) {
    double cv[FUP_UNROLL];
    load_cols8(cv, cols);
#pragma unroll
    for (int u = 0; u < FUP_UNROLL; ++u) {
        t[u].x = cell_update_raw(t[u].x, py.x, py.y, rj.x, cv[u]);
        t[u].y = cell_update_raw(t[u].y, py.x, py.y, rj.y, cv[u]);
    }
}

// the rare road: a batch that holds a pivot row / column, failed the range test or is ragged — the reference's
// formulas with the fully guarded division (pivot row :156, pivot column :160, pivot cell :163, ordinary cells
// :173-175), one row pair per call.  Not inlined and by value: neither its registers nor an addressable copy of the
// batch may weigh on the steady state.
__device__ __noinline__ double2 guarded_pair(double2 t, int f, const LevelDiv *s_lvl, const double2 *my_rows, int rstride,
                                             const double *col, int cstride, int row, int j) {
    for (int l = 0; l < f; ++l) {
        const double2 rj = my_rows[l * rstride];
        const LevelDiv L = s_lvl[l];
        const double ci = col[l * cstride];
        t.x = apply_level(t.x, row, j, L, rj.x, ci);
        t.y = apply_level(t.y, row, j + 1, L, rj.y, ci);
    }
    return t;
}

// WARP-AUTONOMOUS STRIP WALKER.  A warp owns a 64-column strip x `rw` rows (rw a multiple of 8) and walks down it
// in batches of 8 rows; after one barrier at kernel start (the levels' divisors, shared by the CTA) warps never
// synchronise with each other again.  Everything a batch needs arrives ASYNCHRONOUSLY while the previous batch
// computes (cp.async, 16 bytes per lane per row):
//   * the cells: lane L copies ITS pair of columns of the next 8 rows into its own 8 shared-memory slots (nobody
//     else reads them), and moves them to registers with eight 128-bit shared loads when the batch starts;
//   * the 8 levels x 8 rows of COL multipliers of the next batch (512 bytes: one 16-byte copy per lane) into a
//     double-buffered warp-wide slot (one __syncwarp per batch makes them visible);
//   * the 8 ROW values of a lane's two columns are loaded once per strip into the lane's own slots.
// So no warp ever waits for HBM in the steady state, with 2-3 warps per scheduler.  How the design got here
// (profiles/r2/): (i) round 1's per-level range guard was 20 % of the dispatch slots — gone (lazy guard above);
// (ii) CTA-wide TMA staging made the warps of all co-resident CTAs move in lock step (prologue barrier + TMA round
// trip together, hot loops together): the fp64 pipe idled ~25 % of the time although 2.6 warps per scheduler were
// ready on average; a persistent CTA variant with double-buffered TMA stages still paid a barrier per tile and a
// static tail; (iii) warp-private slices without prefetch left ~16 % of the warp time in long-scoreboard waits.
// The dispatch port is the binding resource: an fp64 warp instruction holds it for 2 cycles, any other for 1
// (tools/fp64_lab.cu reproduces every measured pipe utilisation with that model), so every instruction that is not
// one of the 6 fp64 issues per cell-level costs fp64 throughput.
// NP = column PAIRS per lane: a warp's strip is 64 * NP columns (lane L owns columns 2L, 2L + 1 of each 64-column
// half).  NP = 2 halves the per-cell cost of everything that is per batch or per level (multiplier loads, address
// arithmetic, the loop) at the price of 2 x the registers and shared memory per warp (12 warps per SM).
template <int NP> struct LazyGeom {
    static constexpr int SC = 64 * NP;                          // columns per strip
    static constexpr int T_BYTES = 8 * NP * 32 * 16;            // cells of one batch: [8 rows][NP][32 lanes] double2
    static constexpr int R_BYTES = FUSE_MAX * NP * 32 * 16;     // ROW values: [FUSE_MAX][NP][32 lanes] double2
    static constexpr int C_BYTES = 2 * FUSE_MAX * 8 * 8;        // COL multipliers: [2][FUSE_MAX][8 rows]
    static constexpr int WARP_BYTES = T_BYTES + R_BYTES + C_BYTES;
};

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// one level on a lane's 8 x NP pairs, no guard
template <int NP>
__device__ __forceinline__ void lazy_level_np(double2 (&t)[FUP_UNROLL][NP], const double2 *rows_l, const double2 py,
                                              const double *cols) {
    double cv[FUP_UNROLL];
    load_cols8(cv, cols);
    double2 rj[NP];
#pragma unroll
    for (int h = 0; h < NP; ++h) rj[h] = rows_l[h * 32];
#pragma unroll
    for (int u = 0; u < FUP_UNROLL; ++u) {
#pragma unroll
        for (int h = 0; h < NP; ++h) {
            t[u][h].x = cell_update_raw(t[u][h].x, py.x, py.y, rj[h].x, cv[u]);
            t[u][h].y = cell_update_raw(t[u][h].y, py.x, py.y, rj[h].y, cv[u]);
        }
    }
}

template <int NP, int THREADS>
__device__ __forceinline__ void lazy_strip(double *A0, double *A1, int n, int m, int64_t ld, int64_t cbd, int64_t col0, int R,
                                           int rw, const PlanHeader *__restrict__ plan, const double *__restrict__ ROWS,
                                           const double *__restrict__ COLS) {
    using G = LazyGeom<NP>;
    const int f = plan->f;
    if (f <= 0) return;
    extern __shared__ __align__(128) unsigned char lz_raw[];
    __shared__ LevelDiv s_lvl[FUSE_MAX];
    __shared__ __align__(16) double2 s_py[FUSE_MAX];             // (p, y) per level: one 128-bit load
    __shared__ const double *s_colp[FUSE_MAX];                   // COL_l plane of each level
    __shared__ int s_guarded;                                    // some pivot of the pass is outside the lazy guard's span
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char *wbase = lz_raw + (size_t)warp * G::WARP_BYTES;
    double2 *my_t = reinterpret_cast<double2 *>(wbase) + lane;                            // row u, half h: my_t[(u * NP + h) * 32]
    double2 *my_rows = reinterpret_cast<double2 *>(wbase + G::T_BYTES) + lane;            // level l, half h: my_rows[(l * NP + h) * 32]
    double *s_colbuf = reinterpret_cast<double *>(wbase + G::T_BYTES + G::R_BYTES);       // [2][FUSE_MAX][8]

    const int src = plan->src;
    const double *__restrict__ Ain = src ? A1 : A0;
    double *__restrict__ Aout = src ? A0 : A1;

    // ---- once per CTA: the levels
    if (threadIdx.x == 0) s_guarded = 0;
    __syncthreads();
    if ((int)threadIdx.x < f) {
        const int l = threadIdx.x;
        const double p = plan->lvl[l].p;
        s_lvl[l].r = plan->lvl[l].r;
        const int64_t cg = (int64_t)plan->lvl[l].c - col0;           // the plan carries GLOBAL column indices
        s_lvl[l].c = (cg >= 0 && cg < m) ? (int)cg : -1;
        const PivotDiv d = pivot_div_prepare(p);
        s_lvl[l].d = d;
        s_py[l] = make_double2(d.p, d.y);
        s_colp[l] = COLS + ((int64_t)l * R + plan->owner[l]) * cbd;
        const int ep = (__double2hiint(p) >> 20) & 0x7ff;
        if (!d.ok || ep < 1023 - LZ_P_SPAN || ep > 1023 + LZ_P_SPAN) s_guarded = 1;
    }
    __syncthreads();

    const int nstrips = (m + G::SC - 1) / G::SC;
    const int nchunks = (n + 1 + rw - 1) / rw;
    const int w = blockIdx.x * (THREADS / 32) + warp;           // < 2^31: checked by the launcher
    if (w >= nstrips * nchunks) return;
    const int chunk = w / nstrips, strip = w - chunk * nstrips;
    const int j0 = strip * G::SC, i0 = chunk * rw;
    const int jj = j0 + 2 * lane;                                // half h: columns jj + 64 h, jj + 64 h + 1
    const int rows = min(rw, n + 1 - i0);
    const int nb = rows >> 3;                                    // whole batches
    bool active[NP];
#pragma unroll
    for (int h = 0; h < NP; ++h) active[h] = jj + 64 * h < m;

    // row pitch in bytes fits 32 bits (ld <= 2^28 doubles): every row address is ONE IMAD.WIDE.U32 off the batch base
    const uint32_t ldb = (uint32_t)(ld * 8);
    const int64_t delta = reinterpret_cast<const char *>(Aout) - reinterpret_cast<const char *>(Ain);
    const char *sp = reinterpret_cast<const char *>(Ain + (int64_t)i0 * ld + jj);         // this lane's first pair, first row of the batch
    // this lane's share of a batch's COL multipliers: 16 bytes = rows 2k, 2k + 1 of level l, lane = 4 l + k
    const double *my_colsrc = (lane < 4 * f) ? s_colp[lane >> 2] + i0 + 2 * (lane & 3) : nullptr;
    double *my_coldst = s_colbuf + (lane >> 2) * 8 + 2 * (lane & 3);

    auto fetch_batch = [&](const char *from, int b) {           // cells + multipliers of batch b, asynchronously
#pragma unroll
        for (int h = 0; h < NP; ++h) {
            if (active[h]) {
#pragma unroll
                for (int u = 0; u < FUP_UNROLL; ++u) cp_async16(my_t + (u * NP + h) * 32, from + (uint64_t)ldb * u + 512 * h);
            }
        }
        if (my_colsrc) cp_async16(my_coldst + (b & 1) * (FUSE_MAX * 8), my_colsrc + 8 * b);
        cp_async_commit();
    };

    if (nb > 0) fetch_batch(sp, 0);
#pragma unroll
    for (int h = 0; h < NP; ++h) {                               // this lane's ROW values, once per strip
        if (active[h]) {
            const double *rp = ROWS + jj + 64 * h;
            for (int l = 0; l < f; ++l) my_rows[(l * NP + h) * 32] = __ldg(reinterpret_cast<const double2 *>(rp + (int64_t)l * ld));
        }
    }
    // colmask: the levels whose pivot COLUMN lies in this strip — the warp then runs the lazy levels and overwrites
    // that one column after each (:160); rowhit: some level's pivot ROW lies in this warp's rows (then every batch
    // looks for it); guarded_all: a pivot outside the guard's span sends every batch down the guarded road
    uint32_t colmask = 0;
    bool rowhit = false;
    for (int l = 0; l < f; ++l) {
        const int cl = s_lvl[l].c, rl = s_lvl[l].r - i0;
        if (cl >= j0 && cl < j0 + G::SC) colmask |= 1u << l;
        rowhit = rowhit || (rl >= 0 && rl < rows);
    }
    const bool guarded_all = s_guarded != 0;

    for (int b = 0; b < nb; ++b) {
        cp_async_wait_all();
        __syncwarp();                                            // batch b: my cells, everybody's multipliers
        double2 t[FUP_UNROLL][NP];
#pragma unroll
        for (int u = 0; u < FUP_UNROLL; ++u)
#pragma unroll
            for (int h = 0; h < NP; ++h) t[u][h] = my_t[(u * NP + h) * 32];
        const double *cols = s_colbuf + (b & 1) * (FUSE_MAX * 8);
        bool redo = guarded_all;
        if (rowhit) {
            for (int l = 0; l < f; ++l) {
                const int rl = s_lvl[l].r - i0 - 8 * b;
                redo = redo || (rl >= 0 && rl < 8);
            }
        }
        if (b + 1 < nb) fetch_batch(sp + (uint64_t)ldb * FUP_UNROLL, b + 1);
        if (!redo) {
            if (colmask) {
                for (int l = 0; l < f; ++l) {
                    lazy_level_np<NP>(t, my_rows + l * NP * 32, s_py[l], cols + l * 8);
                    const int cl = s_lvl[l].c - jj;              // 0 / 1 (+ 64 h): this lane holds the pivot column
                    if (((colmask >> l) & 1u) && cl >= 0 && (cl & 63) < 2 && cl < G::SC) {
                        const PivotDiv d = s_lvl[l].d;
#pragma unroll
                        for (int u = 0; u < FUP_UNROLL; ++u) {
                            const double qv = pivot_div(cols[l * 8 + u], d);             // :159-160
#pragma unroll
                            for (int h = 0; h < NP; ++h) {
                                if (cl == 64 * h) t[u][h].x = qv;
                                if (cl == 64 * h + 1) t[u][h].y = qv;
                            }
                        }
                    }
                }
            } else if (f == FUSE_MAX) {
#pragma unroll
                for (int l = 0; l < FUSE_MAX; ++l) lazy_level_np<NP>(t, my_rows + l * NP * 32, s_py[l], cols + l * 8);
            } else {
                for (int l = 0; l < f; ++l) lazy_level_np<NP>(t, my_rows + l * NP * 32, s_py[l], cols + l * 8);
            }
            float lo = __uint_as_float(0x7f000000u), hi = 0.0f;
#pragma unroll
            for (int u = 0; u < FUP_UNROLL; ++u)
#pragma unroll
                for (int h = 0; h < NP; ++h)
                    if (active[h]) range_fold(lo, hi, t[u][h].x, t[u][h].y);   // a half beyond the last column holds stale data
            redo = !range_ok(lo, hi);
            if (__builtin_expect(redo, 0)) {
#pragma unroll
                for (int u = 0; u < FUP_UNROLL; ++u)
#pragma unroll
                    for (int h = 0; h < NP; ++h)
                        if (active[h]) t[u][h] = ld_stream(reinterpret_cast<const double *>(sp + (uint64_t)ldb * u + 512 * h));
            }
        }
        if (__builtin_expect(redo, 0)) {
#pragma unroll
            for (int u = 0; u < FUP_UNROLL; ++u)
#pragma unroll
                for (int h = 0; h < NP; ++h)
                    if (active[h])
                        t[u][h] = guarded_pair(t[u][h], f, s_lvl, my_rows + h * 32, NP * 32, cols + u, 8, i0 + 8 * b + u, jj + 64 * h);
        }
        char *dp = const_cast<char *>(sp) + delta;
#pragma unroll
        for (int h = 0; h < NP; ++h) {
            if (active[h]) {
#pragma unroll
                for (int u = 0; u < FUP_UNROLL; ++u) st_stream(reinterpret_cast<double *>(dp + (uint64_t)ldb * u + 512 * h), t[u][h]);
            }
        }
        sp += (uint64_t)ldb * FUP_UNROLL;
    }
    if (rows & 7) {
        // the ragged last batch of the table (n + 1 is rarely a multiple of 8): guarded road, multipliers from global
        const int left = rows & 7, row0 = i0 + 8 * nb;
        for (int h = 0; h < NP; ++h) {
            if (!active[h]) continue;
            for (int u = 0; u < left; ++u) {
                double2 t = ld_stream(reinterpret_cast<const double *>(sp + (uint64_t)ldb * u + 512 * h));
                for (int l = 0; l < f; ++l) {
                    const double2 rj = my_rows[(l * NP + h) * 32];
                    const LevelDiv L = s_lvl[l];
                    const double ci = __ldcg(s_colp[l] + row0 + u);
                    t.x = apply_level(t.x, row0 + u, jj + 64 * h, L, rj.x, ci);
                    t.y = apply_level(t.y, row0 + u, jj + 64 * h + 1, L, rj.y, ci);
                }
                st_stream(reinterpret_cast<double *>(const_cast<char *>(sp) + delta + (uint64_t)ldb * u + 512 * h), t);
            }
        }
    }
}

// SYNC = false: the update of one pass, launched after its pricing kernel in stream order.
// SYNC = true : the update of pass `seq` next to a PERSISTENT pricing kernel (fused_run, price engine 2): the kernel is
//   launched ahead of time; every CTA waits until the pricing kernel has published the pass (plan, ROW planes, COL
//   planes: cs->plan_ready >= seq, acquire) and the last CTA to finish publishes cs->upd_done = seq (release), which
//   the pricing of pass seq + 2 waits for before it gathers from the table written here.
template <int NP, int THREADS, int MINB, bool SYNC>
__global__ void __launch_bounds__(THREADS, MINB)
update_lazy_kernel(double *A0, double *A1, int n, int m, int64_t ld, int64_t cbd, int64_t col0, int R, int rw,
                   const PlanHeader *__restrict__ plan, const double *__restrict__ ROWS,
                   const double *__restrict__ COLS, CoopScratch *cs, unsigned long long seq) {
    bool go = true;
    if (SYNC) go = cta_wait_ge(&cs->plan_ready, seq, &cs->abort);
    if (go) lazy_strip<NP, THREADS>(A0, A1, n, m, ld, cbd, col0, R, rw, plan, ROWS, COLS);
    if (SYNC) {
        __syncthreads();                                         // every warp's stores are issued
        if (threadIdx.x == 0) {
            __threadfence();
            const unsigned int t = atomicAdd(&cs->upd_ctr, 1u);
            if (t == gridDim.x - 1) {
                cs->upd_ctr = 0u;                                // the next update kernel starts after this one
                __threadfence();
                st_release_gpu_u64(&cs->upd_done, seq);
            }
        }
    }
}

// adversarial self-test of the lazy guard: chain k of `count` starts from cell t[k] and goes through F levels
// (p[l], ROW value rj[k][l], COL value ci[k][l]); out_lazy = the kernel's policy (raw chain, range test on the
// output, guarded redo when it fails), out_ref = the guarded chain.  The host compares bits.
__global__ void lazy_guard_selftest_kernel(const double *__restrict__ t0, const double *__restrict__ p,
                                           const double *__restrict__ rj, const double *__restrict__ ci, int F,
                                           int64_t count, double *__restrict__ out_lazy, double *__restrict__ out_ref,
                                           unsigned long long *__restrict__ n_redo) {
    __shared__ PivotDiv s_d[FUSE_MAX];
    __shared__ int s_guarded;
    if (threadIdx.x == 0) s_guarded = 0;
    __syncthreads();
    if ((int)threadIdx.x < F) {
        const PivotDiv d = pivot_div_prepare(p[threadIdx.x]);
        s_d[threadIdx.x] = d;
        const int ep = (__double2hiint(d.p) >> 20) & 0x7ff;
        if (!d.ok || ep < 1023 - LZ_P_SPAN || ep > 1023 + LZ_P_SPAN) s_guarded = 1;
    }
    __syncthreads();
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < count; k += (int64_t)gridDim.x * blockDim.x) {
        double ref = t0[k], lz = t0[k];
        for (int l = 0; l < F; ++l) ref = cell_update(ref, s_d[l], rj[k * F + l], ci[k * F + l]);
        bool redo = s_guarded != 0;
        if (!redo) {
            for (int l = 0; l < F; ++l) lz = cell_update_raw(lz, s_d[l].p, s_d[l].y, rj[k * F + l], ci[k * F + l]);
            float lo = __uint_as_float(0x7f000000u), hi = 0.0f;
            range_fold(lo, hi, lz, lz);
            redo = !range_ok(lo, hi);
        }
        if (redo) {
            atomicAdd(n_redo, 1ull);
            lz = t0[k];
            for (int l = 0; l < F; ++l) lz = cell_update(lz, s_d[l], rj[k * F + l], ci[k * F + l]);
        }
        out_lazy[k] = lz;
        out_ref[k] = ref;
    }
}

} // namespace

namespace spx_launch {

int64_t colbuf_doubles(int n);
int sm_count();

int fuse_max() { return FUSE_MAX; }

static inline int64_t align128(int64_t v) { return (v + 127) / 128 * 128; }

// workspace: plan[2] | ROWS[2][FUSE_MAX][ld] | COLS[FUSE_MAX][cbd] | frow[ld] | bvec[n] | bvec2[n] | CoopScratch |
//            ratio partials [FUSE_MAX][COOP_MAX_CTAS] | a one-rank XBOX (the look-ahead loop on ONE GPU runs
//            the sharded pricing kernel with R = 1).  plan/ROWS are double-buffered by pass parity: the
//            pricing of pass q+1 writes one set while the update of pass q reads the other.
constexpr int COOP_MAX_CTAS = 160;

struct FusedWork {
    PlanHeader *plan[2];
    double *ROWS[2];
    double *COLS, *frow, *bvec, *bvec2;
    CoopScratch *cs;
    Ratio *part;
    unsigned char *xbox1;
    int64_t bytes;
};

FusedWork carve_work(void *work, int n, int64_t ld) {
    const int64_t cbd = colbuf_doubles(n);
    char *p = static_cast<char *>(work);
    char *p0 = p;
    FusedWork w;
    for (int h = 0; h < 2; ++h) { w.plan[h] = reinterpret_cast<PlanHeader *>(p); p += align128(sizeof(PlanHeader)); }
    for (int h = 0; h < 2; ++h) { w.ROWS[h] = reinterpret_cast<double *>(p); p += align128(FUSE_MAX * ld * 8); }
    w.COLS = reinterpret_cast<double *>(p);   p += align128(FUSE_MAX * cbd * 8);
    w.frow = reinterpret_cast<double *>(p);   p += align128(ld * 8);
    w.bvec = reinterpret_cast<double *>(p);   p += align128(((int64_t)n + 16) * 8);
    w.bvec2 = reinterpret_cast<double *>(p);  p += align128(((int64_t)n + 16) * 8);
    w.cs = reinterpret_cast<CoopScratch *>(p); p += align128(sizeof(CoopScratch));
    w.part = reinterpret_cast<Ratio *>(p);    p += align128((int64_t)FUSE_MAX * COOP_MAX_CTAS * sizeof(Ratio));
    w.xbox1 = reinterpret_cast<unsigned char *>(p); p += align128(xbox_layout(cbd, 1).bytes);
    w.bytes = p - p0;
    return w;
}

int64_t fused_workspace_bytes(int n, int64_t ld) { return carve_work(nullptr, n, ld).bytes; }

int64_t get_option(int key);

// One-time per-device configuration of every update_lazy_kernel instantiation (dynamic shared-memory limit).  It also
// LOADS them: with CUDA's lazy module loading the first use of a kernel may need a context-wide synchronisation, which
// never returns while a persistent kernel that waits for that very kernel is resident (see preload_engine_kernels).
static cudaError_t configure_lazy_kernels() {
    static bool configured[64] = {};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
    if (configured[dev]) return cudaSuccess;
    cudaFuncAttributes fa;
#define SPX_LZ_CFG1(NP, TH, MB, SY) \
    if ((e = cudaFuncSetAttribute(update_lazy_kernel<NP, TH, MB, SY>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                  (TH / 32) * LazyGeom<NP>::WARP_BYTES)) != cudaSuccess) return e; \
    if ((e = cudaFuncGetAttributes(&fa, update_lazy_kernel<NP, TH, MB, SY>)) != cudaSuccess) return e;
#define SPX_LZ_CFG(NP, TH, MB) SPX_LZ_CFG1(NP, TH, MB, false) SPX_LZ_CFG1(NP, TH, MB, true)
    SPX_LZ_CFG(1, 256, 2) SPX_LZ_CFG(1, 256, 3) SPX_LZ_CFG(2, 128, 2) SPX_LZ_CFG(2, 128, 3)
#undef SPX_LZ_CFG
#undef SPX_LZ_CFG1
    configured[dev] = true;
    return cudaSuccess;
}

// The fused update launch.  SPX_OPT_FUSE_VARIANT 0 (default): update_lazy_kernel; 1: round 1's update_fused_kernel
// (returns cudaErrorNotSupported here and the caller launches it).  SPX_OPT_FUSE_TILE_ROWS: tile height of the lazy
// kernel (0 = 32; a multiple of 8 <= 256); minb: resident CTAs per SM its register budget targets (0 = 3).
static cudaError_t launch_update_variant(double *A0, double *A1, int n, int m, int64_t ld, int64_t cbd, int64_t col0, int R,
                                         const PlanHeader *plan, const double *ROWS, const double *COLS, int minb,
                                         CoopScratch *sync, unsigned long long seq, cudaStream_t stream) {
    if ((int)get_option(SPX_OPT_FUSE_VARIANT) == 1) return cudaErrorNotSupported;
    int rw = (int)get_option(SPX_OPT_FUSE_TILE_ROWS);
    if (rw <= 0) rw = 128;
    int np = (int)get_option(SPX_OPT_FUSE_PAIRS);
    if (np <= 0) np = 2;
    if (minb <= 0) minb = np == 2 ? 3 : 2;
    cudaError_t e = configure_lazy_kernels();
    if (e != cudaSuccess) return e;
    const int sc = 64 * np, wpc = (np == 2) ? 4 : 8;        // strip width, warps per CTA
    const int64_t nwork = (int64_t)((m + sc - 1) / sc) * ((n + 1 + rw - 1) / rw);      // warp-sized pieces
    if (m <= 0 || nwork == 0) return cudaSuccess;           // a shard without columns only prices
    if (nwork + 64 >= (int64_t)1 << 31) return cudaErrorInvalidValue;
    const unsigned grid = (unsigned)((nwork + wpc - 1) / wpc);
#define SPX_LZ_RUN1(NP, TH, MB, SY) \
    update_lazy_kernel<NP, TH, MB, SY><<<grid, TH, (TH / 32) * LazyGeom<NP>::WARP_BYTES, stream>>>(A0, A1, n, m, ld, cbd, col0, R, \
                                                                                                   rw, plan, ROWS, COLS, sync, seq)
#define SPX_LZ_RUN(NP, TH, MB) do { if (sync) SPX_LZ_RUN1(NP, TH, MB, true); else SPX_LZ_RUN1(NP, TH, MB, false); } while (0)
    if (np == 2) { if (minb == 2) SPX_LZ_RUN(2, 128, 2); else SPX_LZ_RUN(2, 128, 3); }
    else         { if (minb == 3) SPX_LZ_RUN(1, 256, 3); else SPX_LZ_RUN(1, 256, 2); }
#undef SPX_LZ_RUN
#undef SPX_LZ_RUN1
    spx_host::count_launch();
    return cudaGetLastError();
}

cudaError_t lazy_guard_selftest(const double *t0, const double *p, const double *rj, const double *ci, int F, int64_t count,
                                double *out_lazy, double *out_ref, unsigned long long *n_redo, cudaStream_t stream) {
    lazy_guard_selftest_kernel<<<592, 256, 0, stream>>>(t0, p, rj, ci, F, count, out_lazy, out_ref, n_redo);
    spx_host::count_launch();
    return cudaGetLastError();
}

static int g_coop_ctas = -1;      // co-resident CTAs of coop_price_kernel (0: cooperative launch unavailable)

// one pass: price up to F pivots from the materialised table, then stream the body once
cudaError_t fused_pass(double *A0, double *A1, double *b0, double *b1, int n, int m, int64_t ld, int rule,
                       int F, int minb, int pricing, int phase, spx_state *st, void *work, int32_t *rowlab,
                       int32_t *collab, int32_t *trace, cudaStream_t stream) {
    if (F < 1) F = 1;
    if (F > FUSE_MAX) F = FUSE_MAX;
    const int64_t cbd = colbuf_doubles(n);
    const FusedWork w = carve_work(work, n, ld);
    PlanHeader *plan = w.plan[0];
    double *ROWS = w.ROWS[0], *COLS = w.COLS, *frow = w.frow, *bvec = w.bvec, *bvec2 = w.bvec2;
    CoopScratch *cs = w.cs;
    Ratio *part = w.part;
    CoopArgs ca;
    PriceArgs &a = ca.a;
    a.A[0] = A0; a.A[1] = A1; a.b[0] = b0; a.b[1] = b1;
    a.n = n; a.m = m; a.ld = ld; a.cbd = cbd; a.rule = rule; a.F = F;
    a.st = st; a.plan = plan; a.ROWS = ROWS; a.COLS = COLS; a.frow = frow; a.bvec = bvec;
    a.rowlab = rowlab; a.collab = collab; a.trace = trace;
    ca.bv[0] = bvec; ca.bv[1] = bvec2; ca.cs = cs; ca.part = part;
    if (g_coop_ctas < 0) {
        int dev = 0, coop = 0, per_sm = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
        g_coop_ctas = 0;
        if (coop && cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, coop_price_kernel, COOP_THREADS, 0) == cudaSuccess &&
            per_sm > 0)
            g_coop_ctas = min(COOP_MAX_CTAS, sm_count());
    }
    const int64_t col0 = 0;
    cudaError_t e = cudaSuccess;
    // phase: 0 price + update, 1 price only, 2 update only (bench.py times the two kernels separately)
    // pricing: 0 auto (whole-GPU cooperative kernel when the vectors are long enough), 1 one CTA, 2 cooperative
    const bool coop = (pricing == 2 || (pricing == 0 && max(n, m) >= 4096)) && g_coop_ctas > 0;
    if (phase == 2) {
        // nothing to price
    } else if (coop) {
        int G = (max(n + 1, (int)ld) + COOP_THREADS - 1) / COOP_THREADS;
        G = G > g_coop_ctas ? g_coop_ctas : (G < 1 ? 1 : G);
        void *args[] = {&ca};
        e = cudaLaunchCooperativeKernel((const void *)coop_price_kernel, dim3(G), dim3(COOP_THREADS), args, 0, stream);
    } else {
        block_price_kernel<<<1, PRICE_THREADS, 0, stream>>>(a);
        e = cudaGetLastError();
    }
    if (e != cudaSuccess) return e;
    if (phase != 2) spx_host::count_launch();
    if (phase == 1) return cudaSuccess;
    if ((e = launch_update_variant(A0, A1, n, m, ld, cbd, col0, 1, plan, ROWS, COLS, minb, nullptr, 0ull, stream)) != cudaErrorNotSupported)
        return e;
    static bool configured_dev[64] = {};
    bool &configured = configured_dev[spx_host::device_slot()];
    if (!configured) {
        if ((e = cudaFuncSetAttribute(update_fused_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FusedSmem))) != cudaSuccess) return e;
        if ((e = cudaFuncSetAttribute(update_fused_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FusedSmem))) != cudaSuccess) return e;
        if ((e = cudaFuncSetAttribute(update_fused_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FusedSmem))) != cudaSuccess) return e;
        configured = true;
    }
    dim3 grid((unsigned)((m + FUP_TC - 1) / FUP_TC), (unsigned)((n + 1 + FUP_TR - 1) / FUP_TR));
    switch (minb) {
    case 2: update_fused_kernel<2><<<grid, FUP_THREADS, sizeof(FusedSmem), stream>>>(A0, A1, n, m, ld, cbd, col0, 1, plan, ROWS, COLS); break;
    case 3: update_fused_kernel<3><<<grid, FUP_THREADS, sizeof(FusedSmem), stream>>>(A0, A1, n, m, ld, cbd, col0, 1, plan, ROWS, COLS); break;
    default: update_fused_kernel<4><<<grid, FUP_THREADS, sizeof(FusedSmem), stream>>>(A0, A1, n, m, ld, cbd, col0, 1, plan, ROWS, COLS); break;
    }
    spx_host::count_launch();
    return cudaGetLastError();
}

static int g_shard_ctas = -1;

// Everything one rank needs to run fused passes (also used with R = 1 by spx_solve on one GPU)
struct FusedCtx {
    double *A[2], *b[2];
    int n, m_loc, rule, rank, R;
    int64_t ld, col0;
    spx_state *st;
    void *work;
    int32_t *rowlab, *collab, *trace;
    void *xbox[XB_MAX_RANKS];
    unsigned long long seq;              // passes issued so far (flags carry it; parity = buffer set)
    cudaStream_t side;                   // look-ahead: the pricing stream (high priority)
    cudaEvent_t ev_start, ev_priced[2], ev_upd[2];
};

cudaError_t fused_ctx_streams(FusedCtx &c) {
    int lo = 0, hi = 0;
    cudaError_t e;
    if ((e = cudaDeviceGetStreamPriorityRange(&lo, &hi)) != cudaSuccess) return e;
    if ((e = cudaStreamCreateWithPriority(&c.side, cudaStreamNonBlocking, hi)) != cudaSuccess) return e;
    cudaEvent_t *evs[5] = {&c.ev_start, &c.ev_priced[0], &c.ev_priced[1], &c.ev_upd[0], &c.ev_upd[1]};
    for (cudaEvent_t *ev : evs)
        if ((e = cudaEventCreateWithFlags(ev, cudaEventDisableTiming)) != cudaSuccess) return e;
    return cudaSuccess;
}

void fused_ctx_destroy(FusedCtx &c) {
    if (!c.side) return;
    cudaStreamSynchronize(c.side);
    cudaEventDestroy(c.ev_start);
    for (int h = 0; h < 2; ++h) { cudaEventDestroy(c.ev_priced[h]); cudaEventDestroy(c.ev_upd[h]); }
    cudaStreamDestroy(c.side);
    c.side = nullptr;
}

// One pricing launch: `npass` passes starting at pass number c.seq (the caller has already advanced it to the
// first pass).  npass == 1, persistent == false: the classic one-kernel-per-pass form.
static cudaError_t launch_shard_price(const FusedCtx &c, const FusedWork &w, int depth, int64_t pivots, int npass,
                                      bool with_prev, bool persistent, cudaStream_t stream) {
    ShardArgs sa;
    PriceArgs &a = sa.ca.a;
    a.A[0] = c.A[0]; a.A[1] = c.A[1]; a.b[0] = c.b[0]; a.b[1] = c.b[1];
    a.n = c.n; a.m = c.m_loc; a.ld = c.ld; a.cbd = colbuf_doubles(c.n); a.rule = c.rule; a.F = depth;
    a.st = c.st; a.plan = nullptr; a.ROWS = nullptr; a.COLS = nullptr; a.frow = w.frow; a.bvec = w.bvec;
    a.rowlab = c.rowlab; a.collab = c.collab; a.trace = c.trace;
    sa.ca.bv[0] = w.bvec; sa.ca.bv[1] = w.bvec2; sa.ca.cs = w.cs; sa.ca.part = w.part;
    sa.rank = c.rank; sa.R = c.R; sa.col0 = c.col0; sa.seq = c.seq;
    for (int h = 0; h < 2; ++h) { sa.plans[h] = w.plan[h]; sa.rows[h] = w.ROWS[h]; }
    sa.npass = npass; sa.depth = depth; sa.pivots = pivots;
    sa.with_prev = with_prev ? 1 : 0;
    sa.persistent = persistent ? 1 : 0;
    sa.wait_updates = (c.m_loc > 0) ? 1 : 0;             // a shard without columns launches no update kernels
    for (int g = 0; g < XB_MAX_RANKS; ++g) sa.xbox[g] = g < c.R ? static_cast<unsigned char *>(c.xbox[g]) : nullptr;
    if (g_shard_ctas < 0) {
        int dev = 0, coop = 0, per_sm = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
        g_shard_ctas = 0;
        if (coop && cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, shard_price_kernel, COOP_THREADS, 0) == cudaSuccess &&
            per_sm > 0)
            g_shard_ctas = min(COOP_MAX_CTAS, sm_count());
    }
    if (g_shard_ctas <= 0) return cudaErrorNotSupported;
    // threads per CTA (one CTA per SM).  Measured at N = 8 (profiles/r2/): 128 / 256 / 384 / 512 threads ->
    // 13.5 / 17.6 / 19.0 / 18.7 k pivots/s.
    int threads = (int)get_option(SPX_OPT_SHARD_THREADS);
    if (threads <= 0) threads = 512;
    int G = (max(c.n + 1, (int)c.ld) + threads - 1) / threads;
    // SPX_OPT_SHARD_CTAS caps the grid (one CTA per SM): the SMs the pricing kernel does NOT occupy keep streaming
    // the update of the previous pass while it runs
    const int cap_ctas = (int)get_option(SPX_OPT_SHARD_CTAS);
    int max_ctas = (cap_ctas > 0 && cap_ctas < g_shard_ctas) ? cap_ctas : g_shard_ctas;
    // a persistent engine keeps its SMs (a 512-thread CTA takes a whole one) while it waits for update kernels: it must
    // leave them most of the GPU, or they could never finish
    if (persistent && max_ctas > g_shard_ctas / 4) max_ctas = g_shard_ctas / 4 > 0 ? g_shard_ctas / 4 : 1;
    G = G > max_ctas ? max_ctas : (G < 1 ? 1 : G);
    void *args[] = {&sa};
    cudaError_t e = cudaLaunchCooperativeKernel((const void *)shard_price_kernel, dim3(G), dim3(threads), args, 0, stream);
    if (e == cudaSuccess) spx_host::count_launch();
    return e;
}

static cudaError_t configure_round1_kernels() {
    static bool configured_dev[64] = {};
    bool &configured = configured_dev[spx_host::device_slot()];
    cudaError_t e;
    if (!configured) {
        if ((e = cudaFuncSetAttribute(update_fused_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FusedSmem))) != cudaSuccess) return e;
        if ((e = cudaFuncSetAttribute(update_fused_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FusedSmem))) != cudaSuccess) return e;
        configured = true;
    }
    return cudaSuccess;
}

// `sync` != null: the update of pass c.seq next to a persistent pricing kernel (update_lazy_kernel<.., true>)
static cudaError_t launch_fused_update(const FusedCtx &c, const FusedWork &w, int h, int minb, CoopScratch *sync,
                                       cudaStream_t stream) {
    const int slot3 = (int)(c.seq % 3ull);               // the COL planes of this pass (c.seq = its pass number)
    cudaError_t e = configure_round1_kernels();
    if (e != cudaSuccess) return e;
    const int64_t cbd = colbuf_doubles(c.n);
    const XBoxLayout XL = xbox_layout(cbd, c.R);
    const double *COLS = reinterpret_cast<const double *>(static_cast<unsigned char *>(c.xbox[c.rank]) + XL.cols_off) +
                         (int64_t)slot3 * FUSE_MAX * c.R * cbd;
    if ((e = launch_update_variant(c.A[0], c.A[1], c.n, c.m_loc, c.ld, cbd, c.col0, c.R, w.plan[h], w.ROWS[h], COLS, minb,
                                   sync, c.seq, stream)) != cudaErrorNotSupported)
        return e;
    if (sync) return cudaErrorNotSupported;                  // round 1's kernel has no device-side hand-shake
    dim3 grid((unsigned)((c.m_loc + FUP_TC - 1) / FUP_TC), (unsigned)((c.n + 1 + FUP_TR - 1) / FUP_TR));
    if (grid.x == 0) return cudaSuccess;                     // a shard without columns only prices
    if (minb == 3)
        update_fused_kernel<3><<<grid, FUP_THREADS, sizeof(FusedSmem), stream>>>(c.A[0], c.A[1], c.n, c.m_loc, c.ld, cbd, c.col0,
                                                                                c.R, w.plan[h], w.ROWS[h], COLS);
    else
        update_fused_kernel<4><<<grid, FUP_THREADS, sizeof(FusedSmem), stream>>>(c.A[0], c.A[1], c.n, c.m_loc, c.ld, cbd, c.col0,
                                                                                c.R, w.plan[h], w.ROWS[h], COLS);
    spx_host::count_launch();
    return cudaGetLastError();
}

// one warp: returns once the persistent pricing kernel of the call that starts at pass `seq0` is resident (or gives
// up after ENGINE_WAIT_NS and raises cs->abort: the update kernels behind it then skip their bodies)
__global__ void wait_engine_kernel(CoopScratch *cs, unsigned long long seq0) {
    if (threadIdx.x == 0) spin_ge_or_abort(&cs->p_alive, seq0, &cs->abort);
}

// Everything the persistent engine runs next to itself must be loaded BEFORE it starts: CUDA loads kernels lazily, and
// loading one may synchronise the context — behind a resident kernel that is waiting for the kernel being loaded.
// (Seen on 2 GPUs in fresh processes: the first update launch blocked on the host until the engine's wait for it
// timed out; in the single-process tests earlier cases had already loaded every kernel.)
static cudaError_t preload_engine_kernels() {
    static bool loaded[64] = {};
    const int dev = spx_host::device_slot();
    if (loaded[dev]) return cudaSuccess;
    cudaError_t e = configure_lazy_kernels();
    if (e != cudaSuccess) return e;
    if ((e = configure_round1_kernels()) != cudaSuccess) return e;
    cudaFuncAttributes fa;
    if ((e = cudaFuncGetAttributes(&fa, wait_engine_kernel)) != cudaSuccess) return e;
    if ((e = cudaFuncGetAttributes(&fa, shard_price_kernel)) != cudaSuccess) return e;
    loaded[dev] = true;
    return cudaSuccess;
}

// Enqueue `pivots` pivots as passes of `depth` (the last one shorter).  `engine`:
//   0  no look-ahead: price q, update q, price q+1, ... on `s`.
//   1  look-ahead, one pricing kernel per pass: the pricing of pass q+1 runs on the side stream WHILE the update
//      of pass q streams on `s`; it gathers from the table the running update reads and replays that pass's
//      levels first (up to 2 x depth - 1 pending levels).  P_q waits for U_{q-2} (its table and its buffer set),
//      U_q waits for P_q — both through events.
//   2  look-ahead with a PERSISTENT pricing engine: ONE cooperative pricing kernel prices every pass of the call
//      and stays on its SMs (33 of 148 at n = 16384) from the first pass to the last; all update kernels are
//      enqueued at once behind it and the same two dependencies go through device flags (cs->plan_ready,
//      cs->upd_done).  It takes one cooperative launch and two events per pass off the pricing chain.  Measured on
//      cfg4 (profiles/r2/r2q_r2r_persistent_engine.md): 8 ranks 19.5 k pivots/s against 19.2-19.3 k with engine 1,
//      2 ranks 6.4 k against 7.0 k (the engine never gives its SMs back to the update) — the launches were not what
//      bounds a pricing-bound pass, the per-level chain is; engine 1 stays the default, this one is opt-in.
//      The update kernels must not start before the engine is resident (they would fill every SM and spin on
//      plan_ready while it cannot be placed): a one-warp kernel on `s` waits for cs->p_alive first.
cudaError_t fused_run(FusedCtx &c, int64_t pivots, int depth, int minb, int engine, cudaStream_t s) {
    if (depth < 1) depth = 1;
    if (depth > FUSE_MAX) depth = FUSE_MAX;
    if (pivots <= 0) return cudaSuccess;
    const FusedWork w = carve_work(c.work, c.n, c.ld);
    cudaError_t e;
    if (engine == 2 && (int)get_option(SPX_OPT_FUSE_VARIANT) == 1) engine = 1;
    if (engine != 0) {
        if ((e = cudaEventRecord(c.ev_start, s)) != cudaSuccess) return e;
        if ((e = cudaStreamWaitEvent(c.side, c.ev_start, 0)) != cudaSuccess) return e;
    }
    if (engine == 2) {
        if ((e = preload_engine_kernels()) != cudaSuccess) return e;
        const int64_t npass64 = (pivots + depth - 1) / depth;
        if (npass64 > (1 << 30)) return cudaErrorInvalidValue;
        const int npass = (int)npass64;
        ++c.seq;                                             // the first pass of this call
        const unsigned long long seq0 = c.seq;
        if ((e = cudaMemsetAsync(&w.cs->abort, 0, sizeof(unsigned int), c.side)) != cudaSuccess) return e;
        if ((e = launch_shard_price(c, w, depth, pivots, npass, false, true, c.side)) != cudaSuccess) return e;
        wait_engine_kernel<<<1, 32, 0, s>>>(w.cs, seq0);
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
        spx_host::count_launch();
        for (int q = 0; q < npass; ++q) {
            if (q > 0) ++c.seq;
            if ((e = launch_fused_update(c, w, (int)(c.seq & 1ull), minb, w.cs, s)) != cudaSuccess) return e;
        }
        // The engine's last writes (state, labels, b) are ordered before whatever follows on `s`.  The event is
        // recorded only NOW, after the update kernels have been submitted: streams share a few hardware work queues
        // (CUDA_DEVICE_MAX_CONNECTIONS, 8 by default; a process that has initialised NCCL owns dozens of streams),
        // a queue is consumed in submission order, and "record after the engine ends" submitted ahead of the update
        // kernels on the same queue would hold them back until the engine ends — while the engine waits for them
        // (first seen on 2 GPUs: every run stopped at the first upd_done wait with SPX_PEER_TIMEOUT).
        if ((e = cudaEventRecord(c.ev_priced[0], c.side)) != cudaSuccess) return e;
        return cudaStreamWaitEvent(s, c.ev_priced[0], 0);
    }
    int64_t left = pivots;
    for (int q = 0; left > 0; ++q) {
        const int F = (int)(left < depth ? left : depth);
        const int h = (int)(++c.seq & 1ull);
        if (engine == 1) {
            if (q >= 2 && (e = cudaStreamWaitEvent(c.side, c.ev_upd[h], 0)) != cudaSuccess) return e;
            if ((e = launch_shard_price(c, w, F, F, 1, q > 0, false, c.side)) != cudaSuccess) return e;
            if ((e = cudaEventRecord(c.ev_priced[h], c.side)) != cudaSuccess) return e;
            if ((e = cudaStreamWaitEvent(s, c.ev_priced[h], 0)) != cudaSuccess) return e;
            if ((e = launch_fused_update(c, w, h, minb, nullptr, s)) != cudaSuccess) return e;
            if ((e = cudaEventRecord(c.ev_upd[h], s)) != cudaSuccess) return e;
        } else {
            if ((e = launch_shard_price(c, w, F, F, 1, false, false, s)) != cudaSuccess) return e;
            if ((e = launch_fused_update(c, w, h, minb, nullptr, s)) != cudaSuccess) return e;
        }
        left -= F;
    }
    return cudaSuccess;
}

// spx_solve's fused loop on ONE GPU: the sharded machinery with a single rank whose XBOX lives in the
// caller's workspace; one cached side stream + event set per device
static FusedCtx g_solo[64];

cudaError_t fused_solve_passes(double *A0, double *A1, double *b0, double *b1, int n, int m, int64_t ld, int rule,
                               spx_state *st, void *work, int32_t *rowlab, int32_t *collab, int32_t *trace,
                               int64_t pivots, int depth, int minb, int engine, cudaStream_t s) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
    FusedCtx &c = g_solo[dev];
    if (!c.side && (e = fused_ctx_streams(c)) != cudaSuccess) return e;
    c.A[0] = A0; c.A[1] = A1; c.b[0] = b0; c.b[1] = b1;
    c.n = n; c.m_loc = m; c.rule = rule; c.rank = 0; c.R = 1; c.ld = ld; c.col0 = 0;
    c.st = st; c.work = work; c.rowlab = rowlab; c.collab = collab; c.trace = trace;
    c.xbox[0] = carve_work(work, n, ld).xbox1;
    return fused_run(c, pivots, depth, minb, engine, s);
}

// the caller must not free or reuse a workspace while side-stream work on it may be pending
cudaError_t fused_solo_sync() {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < 64 && g_solo[dev].side) return cudaStreamSynchronize(g_solo[dev].side);
    return cudaSuccess;
}

// debug: the %globaltimer stamps of the LAST shard_price_kernel launch on this workspace, [FUSE_MAX + 1][STAMPS]
cudaError_t fused_debug_stamps(const void *work, int n, int64_t ld, unsigned long long *h_out, cudaStream_t s) {
    const FusedWork w = carve_work(const_cast<void *>(work), n, ld);
    cudaError_t e = cudaMemcpyAsync(h_out, w.cs->stamp, sizeof(w.cs->stamp), cudaMemcpyDeviceToHost, s);
    if (e != cudaSuccess) return e;
    return cudaStreamSynchronize(s);
}
int fused_debug_stamp_count() { return (FUSE_MAX + 1) * STAMPS; }

int64_t xbox_bytes(int n, int R) { return xbox_layout(colbuf_doubles(n), R).bytes; }
int xbox_max_ranks() { return XB_MAX_RANKS; }

} // namespace spx_launch

// ---- C ABI of the column-sharded fused loop (declared in include/spx_b200.h) -------------------
struct spx_fshard {
    spx_launch::FusedCtx c;
    int engine;          // 0 no look-ahead, 1 look-ahead with one pricing kernel per pass, 2 persistent pricing engine
};

extern "C" {

int64_t spx_fshard_xbox_bytes(int32_t n, int32_t nranks) {
    if (n < 1 || nranks < 1 || nranks > spx_launch::xbox_max_ranks()) return -1;
    return spx_launch::xbox_bytes(n, nranks);
}

int spx_fshard_open(spx_fshard **out, int32_t rank, int32_t nranks, int32_t n, int32_t m_loc, int64_t ld_loc,
                    int64_t col0, int32_t rule, double *d_A0, double *d_A1, double *d_b0, double *d_b1,
                    spx_state *d_state, void *d_work, int64_t work_bytes, int32_t *d_rowlab, int32_t *d_collab,
                    int32_t *d_trace, void *const *xboxes) {
    using spx_host::set_error;
    if (!out || !d_A0 || !d_A1 || !d_b0 || !d_b1 || !d_state || !d_work || !d_rowlab || !d_collab || !xboxes ||
        nranks < 1 || nranks > spx_launch::xbox_max_ranks() || rank < 0 || rank >= nranks || n < 1 || m_loc < 0 ||
        ld_loc < 16 || ld_loc % 16 || ld_loc < m_loc || col0 < 0 || ((uintptr_t)d_work & 127) ||
        work_bytes < spx_launch::fused_workspace_bytes(n, ld_loc) ||
        (rule != SPX_RULE_REFERENCE && rule != SPX_RULE_DANTZIG)) {
        set_error("spx_fshard_open: bad arguments");
        return -2;
    }
    spx_fshard *h = new (std::nothrow) spx_fshard();
    if (!h) { set_error("spx_fshard_open: out of host memory"); return -2; }
    spx_launch::FusedCtx &c = h->c;
    c.rank = rank; c.R = nranks; c.n = n; c.m_loc = m_loc; c.rule = rule; c.ld = ld_loc; c.col0 = col0;
    c.A[0] = d_A0; c.A[1] = d_A1; c.b[0] = d_b0; c.b[1] = d_b1;
    c.st = d_state; c.work = d_work; c.rowlab = d_rowlab; c.collab = d_collab; c.trace = d_trace;
    for (int g = 0; g < nranks; ++g) {
        if (!xboxes[g]) { delete h; set_error("spx_fshard_open: null xbox %d", g); return -2; }
        c.xbox[g] = xboxes[g];
    }
    c.seq = 0;
    c.side = nullptr;
    h->engine = 1;
    if (spx_host::check(spx_launch::fused_ctx_streams(c), "side stream")) { delete h; return -1; }
    *out = h;
    return 0;
}

int spx_fshard_set_lookahead(spx_fshard *h, int32_t on) {
    if (!h || on < 0 || on > 2) return -2;
    h->engine = on;
    return 0;
}

// Enqueue `pivots` pivots as passes of `depth` (the last one shorter).  Every rank must make the
// same call.  Asynchronous.
int spx_fshard_enqueue(spx_fshard *h, int64_t pivots, int32_t depth, void *stream) {
    if (!h || pivots < 0) { spx_host::set_error("spx_fshard_enqueue: bad arguments"); return -2; }
    if (depth <= 0) depth = 8;
    return spx_host::check(spx_launch::fused_run(h->c, pivots, depth, 0, h->engine,
                                                 reinterpret_cast<cudaStream_t>(stream)), "fused shard passes");
}

int spx_fshard_read(spx_fshard *h, spx_state *h_state, int32_t *cur_buffer, void *stream) {
    if (!h || !h_state) { spx_host::set_error("spx_fshard_read: bad arguments"); return -2; }
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    if (spx_host::check(cudaStreamSynchronize(s), "sync")) return -1;
    if (spx_host::check(cudaStreamSynchronize(h->c.side), "sync side")) return -1;
    if (spx_host::check(cudaMemcpyAsync(h_state, h->c.st, sizeof(spx_state), cudaMemcpyDeviceToHost, s), "read state")) return -1;
    if (spx_host::check(cudaStreamSynchronize(s), "sync")) return -1;
    if (cur_buffer) *cur_buffer = (int32_t)(h_state->reserved[0] & 1);
    return 0;
}

int spx_fused_debug_stamps(const void *d_work, int32_t n, int64_t ld, uint64_t *h_out, int32_t capacity, void *stream) {
    if (!d_work || !h_out || n < 1 || ld < 16 || capacity < spx_launch::fused_debug_stamp_count()) {
        spx_host::set_error("spx_fused_debug_stamps: bad arguments");
        return -2;
    }
    return spx_host::check(spx_launch::fused_debug_stamps(d_work, n, ld, reinterpret_cast<unsigned long long *>(h_out),
                                                          reinterpret_cast<cudaStream_t>(stream)), "debug stamps");
}

int spx_fshard_close(spx_fshard *h) {
    if (!h) return 0;
    spx_launch::fused_ctx_destroy(h->c);
    delete h;
    return 0;
}

} // extern "C"
