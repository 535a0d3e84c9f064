// spx_api.cu — the extern "C" boundary of libspx_b200.so (see include/spx_b200.h).
// Plain pointers and sizes only; no torch types.  Every entry validates its
// arguments, launches on the caller's stream and reports failures through
// spx_last_error().
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "spx_common.cuh"

namespace spx_launch {
int     sm_count();
int64_t colbuf_doubles(int n);
int64_t shard_msg_doubles(int n);
int64_t batched_max_cells();
cudaError_t pick(const double *, const double *, int, int, int64_t, int, int, spx_state *, double *,
                 cudaStream_t);
cudaError_t shard_candidate(const double *, const double *, int, int, int64_t, int64_t, int, int,
                            const spx_state *, double *, cudaStream_t);
cudaError_t shard_select(const double *, int, const double *, int, int, spx_state *, double *,
                         const unsigned long long *, unsigned long long, cudaStream_t);
cudaError_t update(const double *, double *, const double *, double *, int, int, int64_t, int64_t,
                   spx_state *, const double *, int32_t *, int32_t *, int32_t *, int, cudaStream_t);
cudaError_t ahead_candidate(const double *, const double *, double *, int, int, int64_t, int64_t, int,
                            const spx_state *, const double *, double *, cudaStream_t);
cudaError_t ahead_select(const double *, int, const double *, int, const spx_state *, spx_state *,
                         double *, const unsigned long long *, unsigned long long, cudaStream_t);
cudaError_t extract(const double *, int, int, const int32_t *, const double *, double *, double *,
                    cudaStream_t);
bool resident_fits(int, int, int64_t);
cudaError_t resident_loop(double *, double *, double *, double *, int, int, int64_t, int, int64_t, spx_state *,
                          double *, int32_t *, int32_t *, int32_t *, cudaStream_t);
int fuse_max();
int64_t fused_workspace_bytes(int, int64_t);
cudaError_t fused_pass(double *, double *, double *, double *, int, int, int64_t, int, int, int, int, int, spx_state *,
                       void *, int32_t *, int32_t *, int32_t *, cudaStream_t);
cudaError_t fused_solve_passes(double *, double *, double *, double *, int, int, int64_t, int, spx_state *, void *,
                               int32_t *, int32_t *, int32_t *, int64_t, int, int, int, cudaStream_t);
cudaError_t fused_solo_sync();
int64_t get_option(int);
int     set_option(int, int64_t);
cudaError_t selftest_division(const double *, const double *, int64_t, int64_t, unsigned long long *,
                              double *, cudaStream_t);
cudaError_t lazy_guard_selftest(const double *, const double *, const double *, const double *, int, int64_t,
                                double *, double *, unsigned long long *, cudaStream_t);
cudaError_t init_state(spx_state *, int32_t *, int32_t *, int, int, int64_t, cudaStream_t);
cudaError_t solve_batched(double *, int64_t, int, int, int, int, double *, double *, int32_t *,
                          int32_t *, int32_t *, int32_t *, int32_t *, double *, cudaStream_t);
} // namespace spx_launch

namespace spx_host {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int check(cudaError_t e, const char *what) {
    if (e == cudaSuccess) return 0;
    set_error("%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
    return -1;
}

void count_launch(int k) { g_launches.fetch_add(k, std::memory_order_relaxed); }

} // namespace spx_host

using spx_host::check;
using spx_host::set_error;

#define SPX_REQUIRE(cond, ...)            \
    do {                                  \
        if (!(cond)) {                    \
            set_error(__VA_ARGS__);       \
            return -2;                    \
        }                                 \
    } while (0)

static inline cudaStream_t as_stream(void *s) { return reinterpret_cast<cudaStream_t>(s); }

extern "C" {

int spx_version(void) { return SPX_ABI_VERSION; }

const char *spx_last_error(void) { return spx_host::g_err; }

int64_t spx_ld(int64_t m) {
    int64_t ld = (m + 15) / 16 * 16;
    return ld < 16 ? 16 : ld;
}

int64_t spx_cells(int32_t n, int32_t m) { return (int64_t)n * (m + 1) + m; }

int64_t spx_colbuf_doubles(int32_t n) { return spx_launch::colbuf_doubles(n); }

int spx_state_bytes(void) { return (int)sizeof(spx_state); }

int spx_device_info(int32_t *sm_count, int32_t *cc_major, int32_t *cc_minor) {
    int dev = 0;
    if (check(cudaGetDevice(&dev), "cudaGetDevice")) return -1;
    int sms = 0, maj = 0, min = 0;
    if (check(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev), "attr")) return -1;
    if (check(cudaDeviceGetAttribute(&maj, cudaDevAttrComputeCapabilityMajor, dev), "attr")) return -1;
    if (check(cudaDeviceGetAttribute(&min, cudaDevAttrComputeCapabilityMinor, dev), "attr")) return -1;
    if (sm_count) *sm_count = sms;
    if (cc_major) *cc_major = maj;
    if (cc_minor) *cc_minor = min;
    return 0;
}

int64_t spx_launch_count(int reset) {
    long long v = spx_host::g_launches.load(std::memory_order_relaxed);
    if (reset) spx_host::g_launches.store(0, std::memory_order_relaxed);
    return v;
}

int spx_set_option(int32_t option, int64_t value) {
    if (spx_launch::set_option(option, value) != 0) {
        set_error("spx_set_option: bad option %d / value %lld", option, (long long)value);
        return -2;
    }
    return 0;
}

int64_t spx_get_option(int32_t option) { return spx_launch::get_option(option); }

int spx_selftest_division(const double *d_a, const double *d_p, int64_t count, int64_t np,
                          uint64_t *h_mismatches, double *h_first_bad, void *stream) {
    SPX_REQUIRE(d_a && d_p && h_mismatches && count >= 0 && np >= 1, "spx_selftest_division: bad arguments");
    cudaStream_t s = as_stream(stream);
    unsigned long long *d_cnt = nullptr;
    double *d_bad = nullptr;
    if (check(cudaMalloc(&d_cnt, sizeof(*d_cnt)), "cudaMalloc")) return -1;
    if (check(cudaMalloc(&d_bad, 2 * sizeof(double)), "cudaMalloc")) { cudaFree(d_cnt); return -1; }
    int rc = 0;
    double bad[2] = {0.0, 0.0};
    unsigned long long cnt = 0;
    if (check(cudaMemsetAsync(d_cnt, 0, sizeof(*d_cnt), s), "memset") ||
        check(cudaMemsetAsync(d_bad, 0, 2 * sizeof(double), s), "memset") ||
        check(spx_launch::selftest_division(d_a, d_p, count, np, d_cnt, d_bad, s), "selftest launch") ||
        check(cudaMemcpyAsync(&cnt, d_cnt, sizeof(cnt), cudaMemcpyDeviceToHost, s), "read") ||
        check(cudaMemcpyAsync(bad, d_bad, sizeof(bad), cudaMemcpyDeviceToHost, s), "read") ||
        check(cudaStreamSynchronize(s), "sync"))
        rc = -1;
    cudaFree(d_cnt);
    cudaFree(d_bad);
    *h_mismatches = cnt;
    if (h_first_bad) { h_first_bad[0] = bad[0]; h_first_bad[1] = bad[1]; }
    return rc;
}

int spx_selftest_lazy_guard(const double *d_t0, const double *d_p, const double *d_rj, const double *d_ci,
                            int32_t levels, int64_t count, double *d_out_lazy, double *d_out_ref,
                            uint64_t *h_redo, void *stream) {
    SPX_REQUIRE(d_t0 && d_p && d_rj && d_ci && d_out_lazy && d_out_ref && h_redo && count >= 0 && levels >= 1 &&
                levels <= spx_launch::fuse_max(), "spx_selftest_lazy_guard: bad arguments");
    cudaStream_t s = as_stream(stream);
    unsigned long long *d_cnt = nullptr, cnt = 0;
    if (check(cudaMalloc(&d_cnt, sizeof(*d_cnt)), "cudaMalloc")) return -1;
    int rc = 0;
    if (check(cudaMemsetAsync(d_cnt, 0, sizeof(*d_cnt), s), "memset") ||
        check(spx_launch::lazy_guard_selftest(d_t0, d_p, d_rj, d_ci, levels, count, d_out_lazy, d_out_ref, d_cnt, s),
              "selftest launch") ||
        check(cudaMemcpyAsync(&cnt, d_cnt, sizeof(cnt), cudaMemcpyDeviceToHost, s), "read") ||
        check(cudaStreamSynchronize(s), "sync"))
        rc = -1;
    cudaFree(d_cnt);
    *h_redo = cnt;
    return rc;
}

// ---- layout conversion ---------------------------------------------------------
int spx_import_shard(const double *src_rows, const double *src_function, double *d_A, double *d_b,
                     int32_t n, int32_t m, int64_t col0, int32_t m_loc, int64_t ld_loc, void *stream) {
    SPX_REQUIRE(src_rows && (src_function || m_loc == 0) && d_A, "spx_import_shard: null pointer");
    // m == 0 is a packed block of a rank that owns no columns: rows of ONE cell, the b column (pitch 8 bytes)
    SPX_REQUIRE(n >= 0 && (m >= 1 || (m == 0 && m_loc == 0)) && m_loc >= 0 && col0 >= 0 && col0 + m_loc <= m,
                "spx_import_shard: bad shape n=%d m=%d col0=%lld m_loc=%d", n, m, (long long)col0, m_loc);
    SPX_REQUIRE(ld_loc >= m_loc && ld_loc % 16 == 0, "spx_import_shard: ld=%lld must be a multiple of 16 >= m_loc",
                (long long)ld_loc);
    cudaStream_t s = as_stream(stream);
    const size_t spitch = (size_t)(m + 1) * sizeof(double);
    const size_t dpitch = (size_t)ld_loc * sizeof(double);
    if (m_loc > 0 && n > 0)
        if (check(cudaMemcpy2DAsync(d_A, dpitch, src_rows + col0, spitch, (size_t)m_loc * sizeof(double),
                                    (size_t)n, cudaMemcpyDefault, s), "import body")) return -1;
    if (m_loc > 0)   // the f row has m cells (simplex.py:39)
        if (check(cudaMemcpyAsync(d_A + (size_t)n * ld_loc, src_function + col0,
                                  (size_t)m_loc * sizeof(double), cudaMemcpyDefault, s), "import f row")) return -1;
    if (ld_loc > m_loc)
        if (check(cudaMemset2DAsync(d_A + m_loc, dpitch, 0, (size_t)(ld_loc - m_loc) * sizeof(double),
                                    (size_t)n + 1, s), "zero padding")) return -1;
    if (d_b && n > 0)
        if (check(cudaMemcpy2DAsync(d_b, sizeof(double), src_rows + m, spitch, sizeof(double), (size_t)n,
                                    cudaMemcpyDefault, s), "import b")) return -1;
    return 0;
}

int spx_import_table(const double *src_rows, const double *src_function, double *d_A, double *d_b,
                     int32_t n, int32_t m, int64_t ld, void *stream) {
    SPX_REQUIRE(d_b, "spx_import_table: null d_b");
    return spx_import_shard(src_rows, src_function, d_A, d_b, n, m, 0, m, ld, stream);
}

int spx_export_table(const double *d_A, const double *d_b, double *dst_rows, double *dst_function,
                     int32_t n, int32_t m, int64_t ld, void *stream) {
    SPX_REQUIRE(d_A && d_b && dst_rows && dst_function, "spx_export_table: null pointer");
    SPX_REQUIRE(n >= 0 && m >= 1 && ld >= m, "spx_export_table: bad shape");
    cudaStream_t s = as_stream(stream);
    const size_t dpitch = (size_t)(m + 1) * sizeof(double);
    const size_t spitch = (size_t)ld * sizeof(double);
    if (n > 0) {
        if (check(cudaMemcpy2DAsync(dst_rows, dpitch, d_A, spitch, (size_t)m * sizeof(double), (size_t)n,
                                    cudaMemcpyDefault, s), "export body")) return -1;
        if (check(cudaMemcpy2DAsync(dst_rows + m, dpitch, d_b, sizeof(double), sizeof(double), (size_t)n,
                                    cudaMemcpyDefault, s), "export b")) return -1;
    }
    if (check(cudaMemcpyAsync(dst_function, d_A + (size_t)n * ld,
                              (size_t)m * sizeof(double), cudaMemcpyDefault, s), "export f row")) return -1;
    return 0;
}

int spx_init_state(spx_state *d_state, int32_t *d_rowlab, int32_t *d_collab, int32_t n, int32_t m,
                   int64_t max_pivots, void *stream) {
    SPX_REQUIRE(d_state && d_rowlab && d_collab, "spx_init_state: null pointer");
    SPX_REQUIRE(n >= 0 && m >= 1 && max_pivots >= 0, "spx_init_state: bad arguments");
    return check(spx_launch::init_state(d_state, d_rowlab, d_collab, n, m, max_pivots, as_stream(stream)),
                 "init_state launch");
}

// ---- K1 + K2 ------------------------------------------------------------------
static int validate_split(const char *who, const void *A, const void *b, int n, int m, int64_t ld) {
    SPX_REQUIRE(A && b, "%s: null tableau pointer", who);
    SPX_REQUIRE(n >= 1 && m >= 0, "%s: bad shape n=%d m=%d", who, n, m);
    SPX_REQUIRE(ld >= m && ld >= 16 && ld % 16 == 0, "%s: ld=%lld must be a multiple of 16 >= m", who, (long long)ld);
    SPX_REQUIRE(((uintptr_t)A & 127) == 0, "%s: tableau body must be 128-byte aligned", who);
    return 0;
}

int spx_pick(const double *d_A, const double *d_b, int32_t n, int32_t m, int64_t ld, int32_t rule,
             int32_t sticky, spx_state *d_state, double *d_colbuf, void *stream) {
    if (validate_split("spx_pick", d_A, d_b, n, m, ld)) return -2;
    SPX_REQUIRE(d_state && d_colbuf, "spx_pick: null state/colbuf");
    SPX_REQUIRE(rule == SPX_RULE_REFERENCE || rule == SPX_RULE_DANTZIG, "spx_pick: unknown rule %d", rule);
    return check(spx_launch::pick(d_A, d_b, n, m, ld, rule, sticky, d_state, d_colbuf, as_stream(stream)),
                 "pick launch");
}

// ---- K3 -----------------------------------------------------------------------
int spx_shard_update(const double *d_Ain, double *d_Aout, const double *d_bin, double *d_bout,
                     int32_t n, int32_t m_loc, int64_t ld_loc, int64_t col0, spx_state *d_state,
                     const double *d_colbuf, int32_t *d_rowlab, int32_t *d_collab, int32_t *d_trace,
                     int32_t ahead, void *stream) {
    if (validate_split("spx_update", d_Ain, d_bin, n, m_loc, ld_loc)) return -2;
    if (validate_split("spx_update", d_Aout, d_bout, n, m_loc, ld_loc)) return -2;
    SPX_REQUIRE(d_Ain != d_Aout && d_bin != d_bout, "spx_update: the pivot is out of place; in == out");
    SPX_REQUIRE(d_state && d_colbuf && d_rowlab && d_collab, "spx_update: null state/colbuf/labels");
    SPX_REQUIRE(((uintptr_t)d_colbuf & 15) == 0, "spx_update: colbuf must be 16-byte aligned");
    return check(spx_launch::update(d_Ain, d_Aout, d_bin, d_bout, n, m_loc, ld_loc, col0, d_state, d_colbuf,
                                    d_rowlab, d_collab, d_trace, ahead ? 1 : 0, as_stream(stream)),
                 "update launch");
}

int spx_update(const double *d_Ain, double *d_Aout, const double *d_bin, double *d_bout, int32_t n,
               int32_t m, int64_t ld, spx_state *d_state, const double *d_colbuf, int32_t *d_rowlab,
               int32_t *d_collab, int32_t *d_trace, void *stream) {
    return spx_shard_update(d_Ain, d_Aout, d_bin, d_bout, n, m, ld, 0, d_state, d_colbuf, d_rowlab,
                            d_collab, d_trace, 0, stream);
}

// ---- the pivot loop -------------------------------------------------------------
namespace {

// look-ahead workspace: [second state | second colbuf | candidate message]
struct Workspace {
    spx_state *state2;
    double    *colbuf2;
    double    *msg;
};
inline int64_t align128(int64_t v) { return (v + 127) / 128 * 128; }
int64_t workspace_bytes(int n) {
    return align128(sizeof(spx_state)) + align128(spx_launch::colbuf_doubles(n) * 8) +
           align128(spx_launch::shard_msg_doubles(n) * 8);
}
Workspace carve(void *base, int n) {
    char *p = static_cast<char *>(base);
    Workspace w;
    w.state2 = reinterpret_cast<spx_state *>(p);  p += align128(sizeof(spx_state));
    w.colbuf2 = reinterpret_cast<double *>(p);    p += align128(spx_launch::colbuf_doubles(n) * 8);
    w.msg = reinterpret_cast<double *>(p);
    return w;
}

// one high-priority side stream + fork/join events per device, created on first use
struct SideCtx { cudaStream_t side = nullptr; cudaEvent_t fork = nullptr, join = nullptr; };
SideCtx g_side[64];

int side_ctx(SideCtx **out) {
    int dev = 0;
    if (check(cudaGetDevice(&dev), "cudaGetDevice")) return -1;
    if (dev < 0 || dev >= 64) { set_error("device index %d out of range", dev); return -1; }
    SideCtx &c = g_side[dev];
    if (!c.side) {
        int lo = 0, hi = 0;
        if (check(cudaDeviceGetStreamPriorityRange(&lo, &hi), "priority range")) return -1;
        if (check(cudaStreamCreateWithPriority(&c.side, cudaStreamNonBlocking, hi), "side stream")) return -1;
        if (check(cudaEventCreateWithFlags(&c.fork, cudaEventDisableTiming), "event")) return -1;
        if (check(cudaEventCreateWithFlags(&c.join, cudaEventDisableTiming), "event")) return -1;
    }
    *out = &c;
    return 0;
}

} // namespace

int64_t spx_solve_workspace_bytes(int32_t n) { return workspace_bytes(n); }

int64_t spx_fused_workspace_bytes(int32_t n, int32_t m) {
    return spx_launch::fused_workspace_bytes(n, spx_ld(m));
}

int spx_solve(double *d_A0, double *d_A1, double *d_b0, double *d_b1, int32_t n, int32_t m, int64_t ld,
              int32_t rule, spx_state *d_state, double *d_colbuf, int32_t *d_rowlab, int32_t *d_collab,
              int32_t *d_trace, int32_t chunk, int64_t stop_after, int32_t mode, void *d_work,
              int64_t work_bytes, int32_t *h_status, int64_t *h_npiv, void *stream) {
    if (validate_split("spx_solve", d_A0, d_b0, n, m, ld)) return -2;
    if (validate_split("spx_solve", d_A1, d_b1, n, m, ld)) return -2;
    SPX_REQUIRE(d_state && d_colbuf && d_rowlab && d_collab, "spx_solve: null state/colbuf/labels");
    SPX_REQUIRE(chunk >= 1, "spx_solve: chunk must be >= 1");
    SPX_REQUIRE(rule == SPX_RULE_REFERENCE || rule == SPX_RULE_DANTZIG, "spx_solve: unknown rule %d", rule);
    SPX_REQUIRE(mode >= SPX_LOOP_AUTO && mode <= SPX_LOOP_FUSED, "spx_solve: unknown mode %d", mode);
    if (mode == SPX_LOOP_AUTO) {
        // L2-resident tableaus: one persistent kernel; big ones: look-ahead streaming; else classic
        if (spx_launch::resident_fits(n, m, ld)) mode = SPX_LOOP_RESIDENT;
        else if (d_work != nullptr && ((uintptr_t)d_work & 127) == 0 &&
                 work_bytes >= spx_launch::fused_workspace_bytes(n, ld) &&
                 (int64_t)(n + 1) * ld * 8 >= (256LL << 20)) mode = SPX_LOOP_FUSED;
        // in between (measured, tools/cfg2_lab.py 2000 4000: classic 32, fused 28, look-ahead 23 us/pivot)
        else if (d_work != nullptr && ((uintptr_t)d_work & 127) == 0 && work_bytes >= workspace_bytes(n))
            mode = SPX_LOOP_LOOKAHEAD;
        else mode = SPX_LOOP_CLASSIC;
    }
    SPX_REQUIRE(mode != SPX_LOOP_RESIDENT || spx_launch::resident_fits(n, m, ld),
                "spx_solve: the tableau does not fit the L2-resident loop (n <= 4095, 2 bodies <= 96 MB, cooperative launch)");
    SPX_REQUIRE(mode != SPX_LOOP_LOOKAHEAD || d_work != nullptr, "spx_solve: look-ahead needs a workspace");
    SPX_REQUIRE(mode != SPX_LOOP_FUSED || (d_work != nullptr && ((uintptr_t)d_work & 127) == 0 &&
                                           work_bytes >= spx_launch::fused_workspace_bytes(n, ld)),
                "spx_solve: the fused loop needs a 128-byte aligned workspace of spx_fused_workspace_bytes(n, m) = %lld bytes",
                (long long)spx_launch::fused_workspace_bytes(n, ld));
    const bool ahead = (mode == SPX_LOOP_LOOKAHEAD);
    SPX_REQUIRE(!ahead || (work_bytes >= workspace_bytes(n) && ((uintptr_t)d_work & 127) == 0),
                "spx_solve: workspace must be 128-byte aligned and >= spx_solve_workspace_bytes(n) = %lld bytes",
                (long long)workspace_bytes(n));
    cudaStream_t s = as_stream(stream);
    double *A[2] = {d_A0, d_A1};
    double *b[2] = {d_b0, d_b1};
    spx_state hs;
    if (check(cudaMemcpyAsync(&hs, d_state, sizeof(hs), cudaMemcpyDeviceToHost, s), "read state")) return -1;
    if (check(cudaStreamSynchronize(s), "sync")) return -1;
    if (hs.status == SPX_CAP && hs.npiv < hs.max_pivots) {   // the caller raised the cap: resume
        const int32_t run = SPX_PIVOT;
        if (check(cudaMemcpyAsync(&d_state->status, &run, sizeof(run), cudaMemcpyHostToDevice, s), "resume")) return -1;
        hs.status = SPX_PIVOT;
    }
    SideCtx *sc = nullptr;
    Workspace w{};
    if (ahead) {
        if (side_ctx(&sc)) return -1;
        w = carve(d_work, n);
    }
    int64_t done = 0;
    if (mode == SPX_LOOP_FUSED) {
        // the fused passes flip the ping-pong buffers once per PASS, not per pivot: the index of the
        // current buffer travels in the device state while this call runs
        const int64_t curbuf = hs.npiv & 1;
        if (check(cudaMemcpyAsync(&d_state->reserved[0], &curbuf, sizeof(curbuf), cudaMemcpyHostToDevice, s), "cur")) return -1;
        hs.reserved[0] = curbuf;
    }
    bool priced = false;      // look-ahead: d_state/d_colbuf already hold the pick of the current table
    while (hs.status == SPX_PIVOT) {
        int64_t k = chunk;
        if (stop_after > 0 && stop_after - done < k) k = stop_after - done;
        if (k <= 0) break;
        const int64_t base = hs.npiv;
        if (mode == SPX_LOOP_FUSED) {
            // passes of F pivots: price F levels from the stored table (one CTA, O((n+m)F^2) work), then
            // ONE stream over the body applies them all — 16 B of HBM traffic per cell per F pivots
            int F = (int)spx_launch::get_option(SPX_OPT_FUSE_DEPTH);
            if (F <= 0) F = 8;
            if (F > spx_launch::fuse_max()) F = spx_launch::fuse_max();
            // look-ahead (option): the pricing of pass q+1 runs on a side stream during the update of pass q.
            // Off by default on ONE GPU: pricing is 4 % of a pass there and the overlap costs as much as it
            // hides (measured 3.14 k vs 3.21 k pivots/s); on by default in the column-sharded loop.
            const int la = (int)spx_launch::get_option(SPX_OPT_FUSE_LOOKAHEAD);      // 1 per-pass kernels, 2 persistent engine
            if (la == 1 || la == 2) {
                if (check(spx_launch::fused_solve_passes(d_A0, d_A1, d_b0, d_b1, n, m, ld, rule, d_state, d_work, d_rowlab,
                                                         d_collab, d_trace, k, F,
                                                         (int)spx_launch::get_option(SPX_OPT_FUSE_MIN_BLOCKS), la, s),
                          "fused pass launch")) return -1;
            } else {
                int64_t left = k;
                while (left > 0) {
                    const int Fp = (int)(left < F ? left : F);
                    if (check(spx_launch::fused_pass(d_A0, d_A1, d_b0, d_b1, n, m, ld, rule, Fp,
                                                     (int)spx_launch::get_option(SPX_OPT_FUSE_MIN_BLOCKS),
                                                     (int)spx_launch::get_option(SPX_OPT_FUSE_PRICING), 0, d_state, d_work,
                                                     d_rowlab, d_collab, d_trace, s), "fused pass launch")) return -1;
                    left -= Fp;
                }
            }
            if (check(cudaMemcpyAsync(&hs, d_state, sizeof(hs), cudaMemcpyDeviceToHost, s), "read state")) return -1;
            if (check(cudaStreamSynchronize(s), "fused passes")) return -1;
            if (check(spx_launch::fused_solo_sync(), "fused side stream")) return -1;
            done += hs.npiv - base;
            continue;
        }
        if (mode == SPX_LOOP_RESIDENT) {
            // one persistent cooperative launch applies up to k pivots (or all of them)
            if (stop_after <= 0) k = 1LL << 40;
            else k = stop_after - done;
            if (check(spx_launch::resident_loop(d_A0, d_A1, d_b0, d_b1, n, m, ld, rule, k, d_state, d_colbuf,
                                                d_rowlab, d_collab, d_trace, s), "resident launch")) return -1;
            if (check(cudaMemcpyAsync(&hs, d_state, sizeof(hs), cudaMemcpyDeviceToHost, s), "read state")) return -1;
            if (check(cudaStreamSynchronize(s), "resident loop")) return -1;
            done += hs.npiv - base;
            if (hs.status == SPX_PIVOT && hs.npiv == base) break;     // defensive: no progress
            continue;
        }
        if (!ahead) {
            // classic: pick k, update k, pick k+1, ... on one stream
            for (int64_t q = 0; q < k; ++q) {
                const int cur = (int)((base + q) & 1);
                if (check(spx_launch::pick(A[cur], b[cur], n, m, ld, rule, 1, d_state, d_colbuf, s), "pick launch")) return -1;
                if (check(spx_launch::update(A[cur], A[cur ^ 1], b[cur], b[cur ^ 1], n, m, ld, 0, d_state, d_colbuf,
                                             d_rowlab, d_collab, d_trace, 0, s), "update launch")) return -1;
            }
        } else {
            // look-ahead: while update q streams on `s`, the side stream prices pivot q+1 from the
            // same (old) table into the other state/colbuf; the two join before update q+1
            spx_state *S[2] = {d_state, w.state2};
            double *C[2] = {d_colbuf, w.colbuf2};
            if (!priced)
                if (check(spx_launch::pick(A[base & 1], b[base & 1], n, m, ld, rule, 1, d_state, d_colbuf, s), "pick launch")) return -1;
            for (int64_t q = 0; q < k; ++q) {
                const int cur = (int)((base + q) & 1), si = (int)(q & 1);
                if (check(cudaEventRecord(sc->fork, s), "fork")) return -1;
                if (check(cudaStreamWaitEvent(sc->side, sc->fork, 0), "fork wait")) return -1;
                if (check(spx_launch::ahead_candidate(A[cur], b[cur], b[cur ^ 1], n, m, ld, 0, rule, S[si], C[si],
                                                      w.msg, sc->side), "candidate launch")) return -1;
                if (check(spx_launch::ahead_select(w.msg, 1, b[cur ^ 1], n, S[si], S[si ^ 1], C[si ^ 1], nullptr, 0,
                                                   sc->side), "select launch")) return -1;
                if (check(cudaEventRecord(sc->join, sc->side), "join")) return -1;
                if (check(spx_launch::update(A[cur], A[cur ^ 1], b[cur], b[cur ^ 1], n, m, ld, 0, S[si], C[si],
                                             d_rowlab, d_collab, d_trace, 1, s), "update launch")) return -1;
                if (check(cudaStreamWaitEvent(s, sc->join, 0), "join wait")) return -1;
            }
            if (k & 1) {   // the newest state/column sit in the workspace: bring them home
                if (check(cudaMemcpyAsync(d_state, w.state2, sizeof(spx_state), cudaMemcpyDeviceToDevice, s), "state copy")) return -1;
                if (check(cudaMemcpyAsync(d_colbuf, w.colbuf2, (size_t)(n + 1) * sizeof(double),
                                          cudaMemcpyDeviceToDevice, s), "colbuf copy")) return -1;
            }
            priced = true;
        }
        if (check(cudaMemcpyAsync(&hs, d_state, sizeof(hs), cudaMemcpyDeviceToHost, s), "read state")) return -1;
        if (check(cudaStreamSynchronize(s), "pivot chunk")) return -1;
        done += k;
    }
    if (mode == SPX_LOOP_FUSED) {
        // restore the library-wide invariant "the current table is in buffer npiv & 1"
        const int cur = (int)(hs.reserved[0] & 1), want = (int)(hs.npiv & 1);
        if (cur != want) {
            if (check(cudaMemcpyAsync(A[want], A[cur], (size_t)(n + 1) * ld * sizeof(double), cudaMemcpyDeviceToDevice, s), "table copy")) return -1;
            if (check(cudaMemcpyAsync(b[want], b[cur], (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice, s), "b copy")) return -1;
        }
        const int64_t zero = 0;
        if (check(cudaMemcpyAsync(&d_state->reserved[0], &zero, sizeof(zero), cudaMemcpyHostToDevice, s), "cur")) return -1;
        if (check(cudaStreamSynchronize(s), "sync")) return -1;
    }
    if (h_status) *h_status = hs.status;
    if (h_npiv) *h_npiv = hs.npiv;
    return 0;
}

// one fused pass, or one of its two kernels (bench.py times them separately)
int spx_fused_pass(double *d_A0, double *d_A1, double *d_b0, double *d_b1, int32_t n, int32_t m, int64_t ld,
                   int32_t rule, int32_t depth, int32_t phase, spx_state *d_state, void *d_work, int64_t work_bytes,
                   int32_t *d_rowlab, int32_t *d_collab, int32_t *d_trace, void *stream) {
    if (validate_split("spx_fused_pass", d_A0, d_b0, n, m, ld)) return -2;
    if (validate_split("spx_fused_pass", d_A1, d_b1, n, m, ld)) return -2;
    SPX_REQUIRE(d_state && d_work && d_rowlab && d_collab && ((uintptr_t)d_work & 127) == 0 &&
                work_bytes >= spx_launch::fused_workspace_bytes(n, ld) && depth >= 1 && depth <= spx_launch::fuse_max() &&
                phase >= 0 && phase <= 2, "spx_fused_pass: bad arguments");
    return check(spx_launch::fused_pass(d_A0, d_A1, d_b0, d_b1, n, m, ld, rule, depth,
                                        (int)spx_launch::get_option(SPX_OPT_FUSE_MIN_BLOCKS),
                                        (int)spx_launch::get_option(SPX_OPT_FUSE_PRICING), phase, d_state, d_work,
                                        d_rowlab, d_collab, d_trace, as_stream(stream)), "fused pass launch");
}

// ---- find_optimum / f -----------------------------------------------------------
int spx_extract(const double *d_b, int32_t n, int32_t m, const int32_t *d_collab,
                const double *d_function, double *d_x, double *d_obj, void *stream) {
    SPX_REQUIRE(d_b && d_collab && d_function && d_x && d_obj, "spx_extract: null pointer");
    SPX_REQUIRE(n >= 1 && m >= 1, "spx_extract: bad shape");
    return check(spx_launch::extract(d_b, n, m, d_collab, d_function, d_x, d_obj, as_stream(stream)),
                 "extract launch");
}

// ---- K4 -------------------------------------------------------------------------
int64_t spx_batched_max_cells(void) { return spx_launch::batched_max_cells(); }

int spx_solve_batched(double *d_T, int64_t B, int32_t n, int32_t m, int32_t rule, int32_t max_pivots,
                      double *d_x, double *d_obj, int32_t *d_status, int32_t *d_npiv, int32_t *d_rowlab,
                      int32_t *d_collab, int32_t *d_trace, double *d_snap, void *stream) {
    SPX_REQUIRE(B >= 0 && n >= 1 && m >= 1 && max_pivots >= 0, "spx_solve_batched: bad arguments");
    SPX_REQUIRE(B == 0 || (d_T && d_status && d_npiv), "spx_solve_batched: null T/status/npiv");
    SPX_REQUIRE(rule == SPX_RULE_REFERENCE || rule == SPX_RULE_DANTZIG, "spx_solve_batched: unknown rule %d", rule);
    SPX_REQUIRE(spx_cells(n, m) <= spx_batched_max_cells(),
                "spx_solve_batched: %lld cells per LP exceed the shared-memory resident limit %lld; use spx_solve",
                (long long)spx_cells(n, m), (long long)spx_batched_max_cells());
    cudaError_t e = spx_launch::solve_batched(d_T, B, n, m, rule, max_pivots, d_x, d_obj, d_status, d_npiv,
                                              d_rowlab, d_collab, d_trace, d_snap, as_stream(stream));
    return check(e, "batched launch");
}

// ---- column-sharded flow ----------------------------------------------------------
int64_t spx_shard_msg_doubles(int32_t n) { return spx_launch::shard_msg_doubles(n); }

int spx_shard_candidate(const double *d_A, const double *d_b, int32_t n, int32_t m_loc, int64_t ld_loc,
                        int64_t col0, int32_t rule, int32_t sticky, spx_state *d_state, double *d_send,
                        void *stream) {
    if (validate_split("spx_shard_candidate", d_A, d_b, n, m_loc, ld_loc)) return -2;
    SPX_REQUIRE(d_state && d_send && col0 >= 0, "spx_shard_candidate: bad arguments");
    SPX_REQUIRE(rule == SPX_RULE_REFERENCE || rule == SPX_RULE_DANTZIG, "spx_shard_candidate: unknown rule %d", rule);
    return check(spx_launch::shard_candidate(d_A, d_b, n, m_loc, ld_loc, col0, rule, sticky, d_state, d_send,
                                             as_stream(stream)), "candidate launch");
}

int spx_shard_select(const double *d_gathered, int32_t nranks, const double *d_b, int32_t n, int32_t rule,
                     int32_t sticky, spx_state *d_state, double *d_colbuf, const uint64_t *d_flags,
                     uint64_t seq, void *stream) {
    (void)rule;
    SPX_REQUIRE(d_gathered && d_b && d_state && d_colbuf && nranks >= 1 && n >= 1,
                "spx_shard_select: bad arguments");
    return check(spx_launch::shard_select(d_gathered, nranks, d_b, n, sticky, d_state, d_colbuf,
                                          reinterpret_cast<const unsigned long long *>(d_flags), seq,
                                          as_stream(stream)), "select launch");
}

int spx_ahead_candidate(const double *d_A, const double *d_bin, double *d_bout, int32_t n, int32_t m_loc,
                        int64_t ld_loc, int64_t col0, int32_t rule, const spx_state *d_state,
                        const double *d_colbuf, double *d_send, void *stream) {
    if (validate_split("spx_ahead_candidate", d_A, d_bin, n, m_loc, ld_loc)) return -2;
    SPX_REQUIRE(d_bout && d_bout != d_bin && d_state && d_colbuf && d_send && col0 >= 0,
                "spx_ahead_candidate: bad arguments");
    SPX_REQUIRE(rule == SPX_RULE_REFERENCE || rule == SPX_RULE_DANTZIG, "spx_ahead_candidate: unknown rule %d", rule);
    return check(spx_launch::ahead_candidate(d_A, d_bin, d_bout, n, m_loc, ld_loc, col0, rule, d_state, d_colbuf,
                                             d_send, as_stream(stream)), "candidate launch");
}

int spx_ahead_select(const double *d_gathered, int32_t nranks, const double *d_bnext, int32_t n,
                     const spx_state *d_state_cur, spx_state *d_state_next, double *d_colbuf_next,
                     const uint64_t *d_flags, uint64_t seq, void *stream) {
    SPX_REQUIRE(d_gathered && d_bnext && d_state_cur && d_state_next && d_colbuf_next && nranks >= 1 && n >= 1 &&
                d_state_cur != d_state_next, "spx_ahead_select: bad arguments");
    return check(spx_launch::ahead_select(d_gathered, nranks, d_bnext, n, d_state_cur, d_state_next,
                                          d_colbuf_next, reinterpret_cast<const unsigned long long *>(d_flags),
                                          seq, as_stream(stream)), "select launch");
}

} // extern "C"
