// spx_update.cu — K3: the Gauss-Jordan / dictionary rank-1 update,
// recalculate_matrix(), /root/reference/src/simplex.py:149-177.
//
// Memory-bound fp64 stream: every body cell is read once (8 B) and written once
// (8 B) — the 16 B/cell/pivot of BASELINE.json — out of place between two
// ping-pong buffers, so there is no hazard and the pivot row/column are
// immutable for the whole pivot.
//
// Mapping: a CTA owns a tile of TR rows x 512 columns; a thread owns two adjacent
// columns (one 128-bit access per row) and walks down the rows, so
//   - every warp access is a fully coalesced 512-byte line group,
//   - the thread's two pivot-row values live in registers for the whole tile,
//   - the column multiplier of a row is one shared-memory broadcast.
// The 512-column slice of the pivot row and the TR-row slice of the gathered
// pivot column are staged into shared memory with two cp.async.bulk (TMA, SASS
// UBLKCP) copies completing on one mbarrier.  8 rows are loaded before the first
// is consumed: 128 B in flight per thread.
//
// Fused into the same pass: the b ('-b') column update, the label swap
// (:152), the pivot trace, and the pricing of the NEXT pivot (first negative
// new f / new b index, min-reduced into the state) so the next pick starts
// without scanning.
#include "spx_common.cuh"

namespace {

using namespace spx;

constexpr int UPD_THREADS = 256;
constexpr int UPD_TC      = 2 * UPD_THREADS;   // columns per tile
constexpr int UPD_TR_MAX  = 64;                // rows per tile (upper bound; multiple of 8)
constexpr int UPD_UNROLL  = 8;

__global__ void __launch_bounds__(UPD_THREADS)
update_kernel(const double *__restrict__ Ain, double *__restrict__ Aout,
              const double *__restrict__ bin, double *__restrict__ bout,
              int n, int m_loc, int64_t ld, int64_t col0, int tr,
              spx_state *st, const double *__restrict__ colbuf,
              int32_t *__restrict__ rowlab, int32_t *__restrict__ collab,
              int32_t *__restrict__ trace) {
    if (st->status != SPX_PIVOT) return;

    __shared__ alignas(128) double s_row[UPD_TC];
    __shared__ alignas(128) double s_col[UPD_TR_MAX];
    __shared__ alignas(8) uint64_t s_bar;

    const int     r    = st->r;
    const int64_t cg   = st->c;
    const double  p    = st->p;
    const int     slot = st->slot;

    const int tid  = threadIdx.x;
    const int j0   = blockIdx.x * UPD_TC;
    const int i0   = blockIdx.y * tr;
    const int rows = min(tr, n + 1 - i0);

    // ---- stage the pivot-row slice and the pivot-column slice (TMA bulk copies)
    const uint32_t row_bytes = (uint32_t)(min((int64_t)UPD_TC, ld - j0) * 8);   // ld % 16 == 0
    const uint32_t col_bytes = (uint32_t)(tr * 8);        // colbuf is padded to a whole tile
    if (tid == 0) {
        mbar_init(&s_bar, 1);
        mbar_fence_init();
        mbar_expect_tx(&s_bar, row_bytes + col_bytes);
        bulk_g2s(s_row, Ain + (int64_t)r * ld + j0, row_bytes, &s_bar);
        bulk_g2s(s_col, colbuf + i0, col_bytes, &s_bar);
    }
    __syncthreads();               // barrier init visible to the waiters
    mbar_wait(&s_bar, 0);

    const int  j      = j0 + 2 * tid;                    // local column of .x
    const bool active = j < m_loc;
    // column c inside this tile?  jc = 0/1 -> this thread's .x/.y is the pivot column
    const int64_t cl   = cg - col0;                      // local index of the pivot column
    const bool has_c   = (cl >= j0) && (cl < j0 + UPD_TC);
    const int  jc      = has_c ? (int)(cl - j) : -1;

    int fneg = SPX_NONE;                                 // first negative new f index in this thread
    if (active) {
        const double2 rj = *reinterpret_cast<const double2 *>(&s_row[2 * tid]);
        const double *src = Ain + (int64_t)i0 * ld + j;
        double *dst = Aout + (int64_t)i0 * ld + j;
        for (int ii = 0; ii < rows; ii += UPD_UNROLL) {
            double2 t[UPD_UNROLL];
#pragma unroll
            for (int u = 0; u < UPD_UNROLL; ++u)
                if (ii + u < rows) t[u] = ld_stream(src + (int64_t)(ii + u) * ld);
#pragma unroll
            for (int u = 0; u < UPD_UNROLL; ++u) {
                if (ii + u < rows) {
                    const int i = i0 + ii + u;
                    double2 o;
                    if (i == r) {                                        // :155-156
                        o.x = pivot_row_update(t[u].x, p);
                        o.y = pivot_row_update(t[u].y, p);
                        if (jc == 0) o.x = pivot_cell_update(p);         // :163
                        if (jc == 1) o.y = pivot_cell_update(p);
                    } else {                                             // :166-175
                        const double ci = s_col[ii + u];
                        o.x = cell_update(t[u].x, p, rj.x, ci);
                        o.y = cell_update(t[u].y, p, rj.y, ci);
                        if (jc == 0) o.x = pivot_col_update(ci, p);      // :159-160
                        if (jc == 1) o.y = pivot_col_update(ci, p);
                    }
                    st_stream(dst + (int64_t)(ii + u) * ld, o);
                    if (i == n) {                                        // new f row: price the next pivot
                        if (o.x < 0.0) fneg = j;
                        else if (o.y < 0.0 && j + 1 < m_loc) fneg = j + 1;
                    }
                }
            }
        }
    }
    if (i0 + rows == n + 1) {                            // this tile holds the f row (CTA-uniform)
        const int w = __reduce_min_sync(0xffffffffu, fneg);
        if ((tid & 31) == 0 && w != SPX_NONE) atomicMin(&st->hint_fneg[slot], w);
    }

    // ---- the '-b' column (replicated when column-sharded): column tile 0 does it
    if (blockIdx.x == 0) {
        int bneg = SPX_NONE;
        if (tid < rows && i0 + tid < n) {
            const int i = i0 + tid;
            const double bi = bin[i];
            const double nb = (i == r) ? pivot_row_update(bi, p)
                                       : cell_update(bi, p, bin[r], s_col[tid]);
            bout[i] = nb;
            if (nb < 0.0) bneg = i;
        }
        if (tid < 64) {                                  // warps 0,1 cover tr <= 64 rows
            const int w = __reduce_min_sync(0xffffffffu, bneg);
            if ((tid & 31) == 0 && w != SPX_NONE) atomicMin(&st->hint_bneg[slot], w);
        }
        // ---- commit: labels (:152), trace, pivot counter — one thread of one CTA
        if (blockIdx.y == 0 && tid == 0) {
            const int64_t k = st->npiv;
            const int32_t tmp = rowlab[cg]; rowlab[cg] = collab[r]; collab[r] = tmp;
            if (trace) { trace[2 * k] = r; trace[2 * k + 1] = (int32_t)cg; }
            st->hint_tag[slot] = k + 1;
            st->npiv = k + 1;
        }
    }
}

// find_optimum()/f(), simplex.py:48-68, for all m variables
__global__ void extract_kernel(const double *__restrict__ b, int n, int m,
                               const int32_t *__restrict__ collab,
                               const double *__restrict__ function,
                               double *__restrict__ x, double *__restrict__ obj) {
    // x is pre-zeroed by the launcher; labels are unique so the scatter is race-free
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int lab = collab[i];
        if (lab >= 0 && lab < m) x[lab] = b[i];
    }
}

__global__ void objective_kernel(int m, const double *__restrict__ function,
                                 const double *__restrict__ x, double *__restrict__ obj) {
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        obj[0] = (m >= 2) ? __dadd_rn(__dmul_rn(function[0], x[0]), __dmul_rn(function[1], x[1])) : 0.0;
        double s = 0.0;
        for (int j = 0; j < m; ++j) s = __dadd_rn(s, __dmul_rn(function[j], x[j]));
        obj[1] = s;
    }
}

__global__ void init_state_kernel(spx_state *st, int32_t *rowlab, int32_t *collab, int n, int m,
                                  int64_t max_pivots) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int stride = gridDim.x * blockDim.x;
    for (int j = t; j < m; j += stride) rowlab[j] = j;          // 'x1'..'xm'  :30
    for (int i = t; i < n; i += stride) collab[i] = m + i;      // 'y1'..'yn'  :31
    if (t == 0) {
        st->status = SPX_PIVOT; st->r = -1; st->c = -1; st->p = 0.0;
        st->npiv = 0; st->max_pivots = max_pivots; st->phase1 = 0; st->slot = 1;
        st->hint_tag[0] = -1; st->hint_tag[1] = -1;
        st->hint_bneg[0] = st->hint_bneg[1] = SPX_NONE;
        st->hint_fneg[0] = st->hint_fneg[1] = SPX_NONE;
        for (int q = 0; q < 6; ++q) st->reserved[q] = 0;
    }
}

int g_sm_count = 0;

} // namespace

namespace spx_launch {

int sm_count() {
    if (g_sm_count == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev);
        if (g_sm_count <= 0) g_sm_count = 148;
    }
    return g_sm_count;
}

int64_t colbuf_doubles(int n) {
    // whole row tiles for any tile height <= UPD_TR_MAX (the TMA copy reads a full tile)
    return ((int64_t)n + 1 + UPD_TR_MAX - 1) / UPD_TR_MAX * UPD_TR_MAX + UPD_TR_MAX;
}

// rows per tile: 64 for big tableaus; smaller (multiple of 8) when that is what it
// takes to put at least ~4 CTAs on every SM of the device
static int pick_tr(int n, int m_loc) {
    const int64_t ctiles = ((int64_t)m_loc + UPD_TC - 1) / UPD_TC;
    const int64_t want = 4LL * sm_count();
    int tr = UPD_TR_MAX;
    while (tr > 8 && ctiles * (((int64_t)n + 1 + tr - 1) / tr) < want) tr -= 8;
    return tr;
}

cudaError_t update(const double *Ain, double *Aout, const double *bin, double *bout, int n,
                   int m_loc, int64_t ld, int64_t col0, spx_state *st, const double *colbuf,
                   int32_t *rowlab, int32_t *collab, int32_t *trace, cudaStream_t stream) {
    const int tr = pick_tr(n, m_loc);
    dim3 grid((unsigned)((m_loc + UPD_TC - 1) / UPD_TC), (unsigned)((n + 1 + tr - 1) / tr));
    if (grid.x == 0) grid.x = 1;      // a shard with no columns still updates b
    update_kernel<<<grid, UPD_THREADS, 0, stream>>>(Ain, Aout, bin, bout, n, m_loc, ld, col0, tr,
                                                    st, colbuf, rowlab, collab, trace);
    spx_host::count_launch();
    return cudaGetLastError();
}

cudaError_t extract(const double *b, int n, int m, const int32_t *collab, const double *function,
                    double *x, double *obj, cudaStream_t stream) {
    cudaError_t e = cudaMemsetAsync(x, 0, (size_t)m * sizeof(double), stream);
    if (e != cudaSuccess) return e;
    const int blocks = max(1, min(1024, (n + 255) / 256));
    extract_kernel<<<blocks, 256, 0, stream>>>(b, n, m, collab, function, x, obj);
    objective_kernel<<<1, 32, 0, stream>>>(m, function, x, obj);
    spx_host::count_launch(2);
    return cudaGetLastError();
}

cudaError_t init_state(spx_state *st, int32_t *rowlab, int32_t *collab, int n, int m,
                       int64_t max_pivots, cudaStream_t stream) {
    const int blocks = max(1, min(1024, (max(n, m) + 255) / 256));
    init_state_kernel<<<blocks, 256, 0, stream>>>(st, rowlab, collab, n, m, max_pivots);
    spx_host::count_launch();
    return cudaGetLastError();
}

} // namespace spx_launch
