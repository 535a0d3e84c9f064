// spx_batched.cu — K4: whole-LP solver for batches of independent small LPs, no communication:
// the loop of get_solution(), /root/reference/src/simplex.py:179-199, with pick_element()
// (:70-141) and recalculate_matrix() (:143-177) inlined.
//
// The tableau of an LP lives in shared memory in the reference's own flat layout (n rows of m+1
// cells, then the f row with m cells) as two ping-pong copies, because the reference pivots out
// of place (:149) and every cell of the new table reads the OLD pivot row and column.
//   warp mode (<= 96 cells: cfg1 14 cells, cfg3 26 cells): one warp per LP, 8 LPs per CTA; every
//       decision is a warp ballot / redux, all 32 lanes take the same branch, __syncwarp only.
//   CTA mode (Klee-Minty n=20: 440 cells, 2^20-1 strictly sequential pivots): one CTA of 2..16 warps
//       per LP; warp 0 prices the pivot, every thread updates one cell (32 cells per warp), two block
//       barriers per pivot.
// The four cell kinds of the pivot differ only in the numerator, so one division by the pivot
// (pivot_div, reciprocal hoisted) serves them all without divergence.
#include "spx_common.cuh"

namespace {

using namespace spx;

struct BatchedArgs {
    double  *T;          // [B][cells] in/out
    int64_t  B;
    int      n, m, rule, max_pivots;
    double  *x;          // [B][m] or null
    double  *obj;        // [B] or null
    int32_t *status;     // [B]
    int32_t *npiv;       // [B]
    int32_t *rowlab;     // [B][m] or null
    int32_t *collab;     // [B][n] or null
    int32_t *trace;      // [B][max_pivots][2] or null
    double  *snap;       // [B][max_pivots+1][cells] or null
};

// first index q in [0, len) with pred(q); warp-uniform result
template <class F>
__device__ __forceinline__ int warp_first_index(int len, int lane, F pred) {
    for (int base = 0; base < len; base += 32) {
        const int q = base + lane;
        const unsigned ball = __ballot_sync(0xffffffffu, q < len && pred(q));
        if (ball) return base + __ffs(ball) - 1;
    }
    return SPX_NONE;
}

// Cross-lane finish of the ratio scan.  Every lane holds the fold of its own rows; the decision of
// ratio_decide() over the union is taken with redux/ballot instead of a shuffle tree of structs.
__device__ __forceinline__ int warp_ratio_decide(const Ratio &q, bool my_elig_nan, int lane) {
    const unsigned full = 0xffffffffu;
    const int elig = __reduce_min_sync(full, q.elig_row);
    if (elig == SPX_NONE) return -1;                                       // first_try still True
    if (__ballot_sync(full, q.elig_row == elig && my_elig_nan)) return elig;   // NaN min_val is never replaced
    const bool has_neg = q.neg_row >= 0;
    if (__ballot_sync(full, has_neg)) {
        // largest negative ratio: max of the order-preserving 64-bit image, 32 bits at a time
        const unsigned long long key = has_neg ? orderable(q.neg_val) : 0ull;
        const unsigned hi = (unsigned)(key >> 32);
        const unsigned mhi = __reduce_max_sync(full, hi);
        const bool c1 = has_neg && hi == mhi;
        const unsigned lo = c1 ? (unsigned)key : 0u;
        const unsigned mlo = __reduce_max_sync(full, lo);
        const bool c2 = c1 && lo == mlo;
        return __reduce_max_sync(full, c2 ? q.neg_row : -1);               // ties -> highest row (:133, '<=')
    }
    const int zero = __reduce_min_sync(full, q.zero_row);
    return (zero != SPX_NONE) ? zero : -1;                                 // first zero ratio, else min_val > 0
}

// K1 + K2 by ONE warp on the shared-memory table `cur`: returns SPX_PIVOT with (r, c) or the
// terminal status.  All 32 lanes take the same branches (ballots / redux only).
__device__ __forceinline__ int warp_pick(const double *cur, int n, int m, int rule, int lane, int &r, int &c) {
    const int w1 = m + 1, fo = n * w1;
    // ---- K1: phase-1 row (:72-76), its first positive cell (:81-85) ...
    const int r1 = warp_first_index(n, lane, [&](int i) { return cur[i * w1 + m] < 0.0; });
    if (r1 != SPX_NONE) {
        c = warp_first_index(m, lane, [&](int j) { return cur[r1 * w1 + j] > 0.0; });
        if (c == SPX_NONE) return SPX_INCORRECT;                                  // :88-89
        r = r1;                                                                   // :91
        return SPX_PIVOT;
    }
    // ---- ... or the entering column from the f row (:94-98)
    if (rule == SPX_RULE_REFERENCE) {
        c = warp_first_index(m, lane, [&](int j) { return cur[fo + j] < 0.0; });
    } else {   // Dantzig: most negative, lowest index on ties
        unsigned long long best = ~0ull;
        for (int j = lane; j < m; j += 32) {
            const double v = cur[fo + j];
            if (v < 0.0) { const unsigned long long k = orderable(v); best = k < best ? k : best; }
        }
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) {
            const unsigned long long o = __shfl_xor_sync(0xffffffffu, best, s);
            best = o < best ? o : best;
        }
        c = (best == ~0ull) ? SPX_NONE
            : warp_first_index(m, lane, [&](int j) {
                  const double v = cur[fo + j];
                  return v < 0.0 && orderable(v) == best; });
    }
    if (c == SPX_NONE) return SPX_OPTIMAL;                                        // :101-103
    // ---- K2: the ratio scan (:107-136): lane-local fold, then ballots / redux across lanes
    Ratio q = ratio_identity();
    bool my_elig_nan = false;                    // is the ratio of MY first eligible row NaN?
    for (int i = lane; i < n; i += 32) {
        const double a_ic = cur[i * w1 + c], b_i = cur[i * w1 + m];
        const bool first = (q.elig_row == SPX_NONE);
        const bool is_nan = ratio_accumulate(q, i, a_ic, b_i);
        if (first && q.elig_row != SPX_NONE) my_elig_nan = is_nan;
    }
    r = warp_ratio_decide(q, my_elig_nan, lane);
    return (r < 0) ? SPX_NOCONV : SPX_PIVOT;                                      // :138-139
}

// CTA == false: one warp per LP, warps_per_cta LPs per CTA, __syncwarp between the phases
//               (cfg1: 14 cells, cfg3: 26 cells — one cell per lane).
// CTA == true : one CTA per LP for tables of hundreds of cells (Klee-Minty n=20: 440 cells, 2^20-1
//               strictly sequential pivots): warp 0 prices the pivot, every warp updates its share
//               of the cells, two block barriers per pivot.
template <bool CTA>
__global__ void __launch_bounds__(CTA ? 512 : 256, CTA ? 1 : 6)
batched_kernel(BatchedArgs a, int warps_per_cta, int warp_doubles) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int s_ctl[4];                               // CTA mode: {status, r, c} of warp 0's pick
    const int n = a.n, m = a.m, w1 = m + 1;
    const int cells = n * w1 + m;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    // CTA-shared lookup: cell -> (row, col), built once (shape is uniform over the batch)
    uint16_t *cell_i = reinterpret_cast<uint16_t *>(smem_raw);
    uint16_t *cell_j = cell_i + cells;
    for (int k = threadIdx.x; k < cells; k += blockDim.x) {
        const int i = k / w1;
        cell_i[k] = (uint16_t)i;
        cell_j[k] = (uint16_t)(k - i * w1);
    }
    __syncthreads();

    const int64_t lp = CTA ? (int64_t)blockIdx.x : (int64_t)blockIdx.x * warps_per_cta + warp;
    if (!CTA && (warp >= warps_per_cta || lp >= a.B)) return;
    const int tl = CTA ? (int)threadIdx.x : lane;          // thread index within the LP's group
    const int nl = CTA ? (int)blockDim.x : 32;             // threads of the group
    auto group_sync = [&]() { if (CTA) __syncthreads(); else __syncwarp(); };

    const size_t lut_bytes = ((size_t)cells * 4 + 15) / 16 * 16;
    double *base = reinterpret_cast<double *>(smem_raw + lut_bytes) + (size_t)(CTA ? 0 : warp) * warp_doubles;
    double *cur = base;
    double *nxt = base + cells;
    double *xs  = base + 2 * cells;                       // [m]
    int32_t *rl = reinterpret_cast<int32_t *>(xs + m);    // [m] header labels
    int32_t *cl = rl + m;                                 // [n] row labels

    double *Tg = a.T + lp * (int64_t)cells;
    for (int k = tl; k < cells; k += nl) cur[k] = Tg[k];
    for (int j = tl; j < m; j += nl) rl[j] = j;           // 'x1'..'xm'  :30
    for (int i = tl; i < n; i += nl) cl[i] = m + i;       // 'y1'..'yn'  :31
    group_sync();
    const int fo = n * w1;                                // offset of the f row
    const double f0 = (m >= 1) ? cur[fo] : 0.0;           // self.function is never mutated (:29,:49)
    const double f1 = (m >= 2) ? cur[fo + 1] : 0.0;

    int npiv = 0, status;
    double *snap = a.snap ? a.snap + lp * (int64_t)(a.max_pivots + 1) * cells : nullptr;
    int32_t *trace = a.trace ? a.trace + lp * (int64_t)a.max_pivots * 2 : nullptr;

    for (;;) {
        if (snap) {
            double *sg = snap + (int64_t)npiv * cells;
            for (int k = tl; k < cells; k += nl) sg[k] = cur[k];
        }
        int r = -1, c = -1;
        if (!CTA) {
            status = warp_pick(cur, n, m, a.rule, lane, r, c);
        } else {
            if (warp == 0) {
                const int st = warp_pick(cur, n, m, a.rule, lane, r, c);
                if (lane == 0) { s_ctl[0] = st; s_ctl[1] = r; s_ctl[2] = c; }
            }
            __syncthreads();
            status = s_ctl[0]; r = s_ctl[1]; c = s_ctl[2];
        }
        if (status != SPX_PIVOT) break;
        if (npiv >= a.max_pivots) { status = SPX_CAP; break; }
        if (trace && tl == 0) { trace[2 * npiv] = r; trace[2 * npiv + 1] = c; }

        // ---- K3: out-of-place pivot (:149-177), all reads from `cur`; the four cell kinds
        // (:156 pivot row, :160 pivot column, :163 pivot cell, :173-175 the rest) differ only in
        // the numerator, so one division by the pivot serves them all without divergence
        const double p = cur[r * w1 + c];
        const PivotDiv d = pivot_div_prepare(p);
        const double *prow = cur + r * w1;
        for (int k = tl; k < cells; k += nl) {
            const int i = cell_i[k], j = cell_j[k];
            const double t = cur[k];
            const double ci = cur[i * w1 + c];
            const bool pr = (i == r), pc = (j == c);
            double num = __dsub_rn(__dmul_rn(t, p), __dmul_rn(prow[j], ci));
            num = pc ? ci : num;
            num = pr ? -t : num;
            num = (pr && pc) ? 1.0 : num;
            nxt[k] = pivot_div(num, d);
        }
        if (tl == 0) { const int32_t t = rl[c]; rl[c] = cl[r]; cl[r] = t; }      // :152
        group_sync();                                      // (CTA: also orders s_ctl against the next pick)
        double *sw = cur; cur = nxt; nxt = sw;
        ++npiv;
    }

    // ---- epilogue: final table, labels, find_optimum()/f() (:48-68)
    for (int k = tl; k < cells; k += nl) Tg[k] = cur[k];
    for (int j = tl; j < m; j += nl) xs[j] = 0.0;
    group_sync();
    for (int i = tl; i < n; i += nl) {
        const int lab = cl[i];
        if (lab < m) xs[lab] = cur[i * w1 + m];
    }
    group_sync();
    if (a.x) for (int j = tl; j < m; j += nl) a.x[lp * m + j] = xs[j];
    if (a.rowlab) for (int j = tl; j < m; j += nl) a.rowlab[lp * m + j] = rl[j];
    if (a.collab) for (int i = tl; i < n; i += nl) a.collab[lp * n + i] = cl[i];
    if (tl == 0) {
        a.status[lp] = status;
        a.npiv[lp] = npiv;
        if (a.obj) a.obj[lp] = (m >= 2) ? __dadd_rn(__dmul_rn(f0, xs[0]), __dmul_rn(f1, xs[1])) : 0.0;
    }
}

size_t warp_bytes(int n, int m) {
    const size_t cells = (size_t)n * (m + 1) + m;
    size_t b = (2 * cells + m) * sizeof(double) + (size_t)(n + m) * sizeof(int32_t);
    return (b + 15) / 16 * 16;
}
size_t lut_bytes(int n, int m) {
    const size_t cells = (size_t)n * (m + 1) + m;
    return (cells * 4 + 15) / 16 * 16;
}
constexpr size_t SMEM_LIMIT = 227 * 1024;
constexpr int64_t CTA_MODE_MIN_CELLS = 96;    // above this one CTA (not one warp) solves an LP
constexpr int64_t CTA_CELLS_PER_WARP = 32;    // CTA mode: cells per warp (2..8 warps)

} // namespace

namespace spx_launch {

int64_t batched_max_cells() { return 10000; }

// returns cudaErrorInvalidValue when one LP does not fit shared memory
cudaError_t solve_batched(double *T, int64_t B, int n, int m, int rule, int max_pivots, double *x,
                          double *obj, int32_t *status, int32_t *npiv, int32_t *rowlab,
                          int32_t *collab, int32_t *trace, double *snap, cudaStream_t stream) {
    if (B <= 0) return cudaSuccess;
    if (n > 65535 || m > 65534) return cudaErrorInvalidValue;
    const size_t wb = warp_bytes(n, m), lb = lut_bytes(n, m);
    if (lb + wb > SMEM_LIMIT) return cudaErrorInvalidValue;
    const int64_t cells = (int64_t)n * (m + 1) + m;
    BatchedArgs a{T, B, n, m, rule, max_pivots, x, obj, status, npiv, rowlab, collab, trace, snap};
    static size_t configured_dev[64][2] = {};
    size_t *configured = configured_dev[spx_host::device_slot()];
    if (cells > CTA_MODE_MIN_CELLS) {
        // one CTA per LP: one cell per thread where possible (measured on Klee-Minty n=20: 32 cells per warp
        // 891 k pivots/s, 64: 730 k, 128: 623 k), 2..16 warps
        int warps = (int)((cells + CTA_CELLS_PER_WARP - 1) / CTA_CELLS_PER_WARP);
        warps = warps < 2 ? 2 : (warps > 16 ? 16 : warps);
        const size_t smem = lb + wb;
        if (smem > 48 * 1024 && smem > configured[1]) {
            cudaError_t e = cudaFuncSetAttribute(batched_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 (int)SMEM_LIMIT);
            if (e != cudaSuccess) return e;
            configured[1] = SMEM_LIMIT;
        }
        batched_kernel<true><<<(unsigned)B, warps * 32, smem, stream>>>(a, 1, (int)(wb / sizeof(double)));
        spx_host::count_launch();
        return cudaGetLastError();
    }
    // warps per CTA: as many as fit ~48 KB (several CTAs per SM), at most 8, at least 1
    int wpc = (int)((48 * 1024 - lb) / wb);
    if (lb >= 48 * 1024) wpc = 0;
    wpc = wpc > 8 ? 8 : wpc;
    if (wpc < 1) wpc = 1;
    if ((int64_t)wpc > B) wpc = (int)B;
    const size_t smem = lb + (size_t)wpc * wb;
    if (smem > 48 * 1024 && smem > configured[0]) {
        cudaError_t e = cudaFuncSetAttribute(batched_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)SMEM_LIMIT);
        if (e != cudaSuccess) return e;
        configured[0] = SMEM_LIMIT;
    }
    const int64_t ctas = (B + wpc - 1) / wpc;
    batched_kernel<false><<<(unsigned)ctas, wpc * 32, smem, stream>>>(a, wpc, (int)(wb / sizeof(double)));
    spx_host::count_launch();
    return cudaGetLastError();
}

} // namespace spx_launch
