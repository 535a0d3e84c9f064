// spx_block.cuh — CTA-level building blocks shared by the pricing kernels (spx_pick.cu) and the
// L2-resident persistent loop (spx_resident.cu): shuffle -> shared memory -> shuffle reductions of
// the reference's first-index searches (simplex.py:73-76, :82-85, :95-98) and of its ratio scan
// (:107-136).  Every reduction is order independent, so any tree gives the sequential answer.
#pragma once
#include "spx_common.cuh"

namespace spx {

constexpr int BLOCK_MAX_WARPS = 32;

struct Scratch {
    int                red_i[BLOCK_MAX_WARPS];
    unsigned long long red_k[BLOCK_MAX_WARPS];
    Ratio              red_q[BLOCK_MAX_WARPS];
    int                out_i;
    unsigned long long out_k;
    Ratio              out_q;
};

__device__ __forceinline__ int block_min_int(int v, Scratch &s) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    v = warp_min_int(v);
    if (lane == 0) s.red_i[warp] = v;
    __syncthreads();
    if (warp == 0) {
        int w = warp_min_int(lane < (int)(blockDim.x >> 5) ? s.red_i[lane] : SPX_NONE);
        if (lane == 0) s.out_i = w;
    }
    __syncthreads();
    const int out = s.out_i;
    __syncthreads();
    return out;
}

// two independent index minima in one pass over shared memory (one barrier pair instead of two
// reductions): used where two first-index searches have no data dependence on each other
__device__ __forceinline__ void block_min_int2(int &a, int &b, Scratch &s) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (int)(blockDim.x >> 5);
    a = warp_min_int(a);
    b = warp_min_int(b);
    if (lane == 0) { s.red_i[warp] = a; s.red_k[warp] = (unsigned long long)(unsigned)b; }
    __syncthreads();
    a = warp_min_int(lane < nw ? s.red_i[lane] : SPX_NONE);
    b = warp_min_int(lane < nw ? (int)(unsigned)s.red_k[lane] : SPX_NONE);
    __syncthreads();                              // red_i / red_k may be reused right away
}

__device__ __forceinline__ unsigned long long warp_min_u64(unsigned long long v) {
#pragma unroll
    for (int sft = 16; sft > 0; sft >>= 1) {
        unsigned long long o = __shfl_xor_sync(0xffffffffu, v, sft);
        v = o < v ? o : v;
    }
    return v;
}

__device__ __forceinline__ unsigned long long block_min_u64(unsigned long long v, Scratch &s) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    v = warp_min_u64(v);
    if (lane == 0) s.red_k[warp] = v;
    __syncthreads();
    if (warp == 0) {
        unsigned long long w = warp_min_u64(lane < (int)(blockDim.x >> 5) ? s.red_k[lane] : ~0ull);
        if (lane == 0) s.out_k = w;
    }
    __syncthreads();
    const unsigned long long out = s.out_k;
    __syncthreads();
    return out;
}

__device__ __forceinline__ Ratio block_ratio_reduce(Ratio q, Scratch &s) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    q = warp_ratio_reduce(q);
    if (lane == 0) s.red_q[warp] = q;
    __syncthreads();
    if (warp == 0) {
        Ratio w = warp_ratio_reduce(lane < (int)(blockDim.x >> 5) ? s.red_q[lane] : ratio_identity());
        if (lane == 0) s.out_q = w;
    }
    __syncthreads();
    const Ratio out = s.out_q;
    __syncthreads();
    return out;
}

struct IsNeg { __device__ bool operator()(double v) const { return v < 0.0; } };   // :74, :96
struct IsPos { __device__ bool operator()(double v) const { return v > 0.0; } };   // :83

// first j in [0, len) with pred(x[j]); chunked so the usual early hit costs one pass
template <class Pred>
__device__ int block_first_index(const double *__restrict__ x, int len, Pred pred, Scratch &s) {
    const int nt = (int)blockDim.x;
    for (int base = 0; base < len; base += nt * 4) {
        int loc = SPX_NONE;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int j = base + u * nt + (int)threadIdx.x;
            if (j < len && pred(x[j])) loc = min(loc, j);
        }
        loc = block_min_int(loc, s);
        if (loc != SPX_NONE) return loc;
    }
    return SPX_NONE;
}

// same search over values produced on the fly: val(j) is the cell j of a row that does not
// exist in memory yet (the look-ahead kernels price the NEXT table from the current one)
template <class Val, class Pred>
__device__ int block_first_index_fn(int len, Val val, Pred pred, Scratch &s) {
    const int nt = (int)blockDim.x;
    for (int base = 0; base < len; base += nt * 4) {
        int loc = SPX_NONE;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int j = base + u * nt + (int)threadIdx.x;
            if (j < len && pred(val(j))) loc = min(loc, j);
        }
        loc = block_min_int(loc, s);
        if (loc != SPX_NONE) return loc;
    }
    return SPX_NONE;
}

// ---- lazy replay (fused loops): a pending pivot "level" applied to ONE cell ------------------
// Level l = (r, c, p) with ROW_l = the pivot row and COL_l = the pivot column of the table the level
// is applied to.  The cell (t, j) of the next table follows from its current value v, ROW_l[j] and
// COL_l[t] with the reference's formulas and roundings (simplex.py:156, :160, :163, :173-175).
struct LevelDiv { int r, c; PivotDiv d; };

__device__ __forceinline__ double apply_level(double v, int t, int j, const LevelDiv &L, double row_j, double col_t) {
    if (t == L.r) return (j == L.c) ? pivot_cell_update(L.d.p) : pivot_div(-v, L.d);      // :163, :156
    return (j == L.c) ? pivot_div(col_t, L.d) : cell_update(v, L.d, row_j, col_t);         // :160, :173-175
}

} // namespace spx
