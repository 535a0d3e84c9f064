// fp64_lab.cu — developer lab: what the fp64 pipe of a (power-capped) B200 really sustains for the
// instruction mix of the fused update (K6), register-resident, no HBM traffic.
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/fp64_lab tools/fp64_lab.cu
//   tools/fp64_lab
//
// Every variant runs "cell-levels": one cell through one pending pivot level,
//   a = RN(RN(t*p) - RN(rj*ci)); q0 = RN(a*y); rem = fma(-p, q0, a); q = fma(y, rem, q0)      (6 fp64 issues)
// 16 cells per thread (8 rows x 2 columns, like the kernel), 8 levels per batch.  Reported: ns per cell-level per
// SM, fp64 pipe utilisation at the clock measured IN the kernel (clock64 vs globaltimer).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int F = 8, UN = 8, TC = 512, THREADS = 256;

struct Lvl { double p, y; unsigned qlo, pad; };

__device__ __forceinline__ unsigned long long gtimer() {
    unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t;
}

// MODE 0: pure DFMA (6 per cell-level), operands in registers
// MODE 1: real arithmetic, operands from shared memory (rj: LDS.128, p/y: LDS.128, ci: 4 x LDS.128), no guard
// MODE 2: MODE 1 + the round-1 guard (LOP3 + 2 ISETP per cell)
// MODE 3: MODE 1 + FMNMX3 min-accumulate guard (one instruction per 2 cells)
// MODE 4: MODE 1 + IMAD/ISETP guard (2 per cell)
// MODE 5: real arithmetic, ALL operands hoisted into registers (no LDS in the loop), no guard
template <int MODE, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) lab_kernel(const double *in, double *out, int iters, unsigned long long *stamps) {
    __shared__ __align__(16) double s_rows[F][TC];
    __shared__ __align__(16) double s_cols[F][UN];
    __shared__ __align__(16) Lvl s_lvl[F];
    const int tid = threadIdx.x;
    for (int i = tid; i < F * TC; i += THREADS) s_rows[i / TC][i % TC] = in[i % 4096] * 0.001;
    if (tid < F * UN) s_cols[tid / UN][tid % UN] = in[tid] * 0.001;
    if (tid < F) { const double p = 1.0 + 0.03125 * tid; s_lvl[tid].p = p; s_lvl[tid].y = 1.0 / p; s_lvl[tid].qlo = 0x03700000u; }
    __syncthreads();
    double2 t[UN];
#pragma unroll
    for (int u = 0; u < UN; ++u) t[u] = make_double2(in[(tid * 16 + 2 * u) & 4095] + 1.0, in[(tid * 16 + 2 * u + 1) & 4095] + 1.0);
    unsigned long long c0 = clock64(), g0 = gtimer();
    bool ok = true;
    float accmin = __int_as_float(0x7f000000);
    for (int it = 0; it < iters; ++it) {
        asm volatile("" ::: "memory");      // the staged slices change per tile in the real kernel: no hoisting
#pragma unroll
        for (int l = 0; l < F; ++l) {
            if (MODE == 0) {
                const double p = s_lvl[l].p, y = s_lvl[l].y;
#pragma unroll
                for (int u = 0; u < UN; ++u) {
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        t[u].x = __fma_rn(t[u].x, p, y); t[u].y = __fma_rn(t[u].y, p, y);
                        t[u].x = __fma_rn(t[u].x, y, p); t[u].y = __fma_rn(t[u].y, y, p);
                    }
                }
            } else {
                const double2 rj = *reinterpret_cast<const double2 *>(&s_rows[l][2 * tid]);
                const double2 py = *reinterpret_cast<const double2 *>(&s_lvl[l].p);
                const double p = py.x, y = py.y;
                const unsigned qlo = s_lvl[l].qlo;
                double cv[UN];
#pragma unroll
                for (int u = 0; u < UN; u += 2) {
                    const double2 c2 = *reinterpret_cast<const double2 *>(&s_cols[l][u]);
                    cv[u] = c2.x; cv[u + 1] = c2.y;
                }
#pragma unroll
                for (int u = 0; u < UN; ++u) {
                    double q[2];
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const double tv = h ? t[u].y : t[u].x;
                        const double r = h ? rj.y : rj.x;
                        const double a = __dsub_rn(__dmul_rn(tv, p), __dmul_rn(r, cv[u]));
                        const double q0 = __dmul_rn(a, y);
                        const double rem = __fma_rn(-p, q0, a);
                        q[h] = __fma_rn(y, rem, q0);
                        if (MODE == 2) {
                            const unsigned hq = (unsigned)__double2hiint(q[h]) & 0x7fffffffu;
                            ok = ok && (hq >= qlo) && (hq <= 0x7f800000u);
                        }
                        if (MODE == 4) {
                            unsigned td;
                            asm("mad.lo.u32 %0, %1, 2, %2;" : "=r"(td) : "r"((unsigned)__double2hiint(q[h])), "r"(0u - 2 * qlo));
                            ok = ok && (td < 0xff000001u - 2 * qlo);
                        }
                    }
                    if (MODE == 3) {
                        const float fx = fabsf(__int_as_float(__double2hiint(q[0])));
                        const float fy = fabsf(__int_as_float(__double2hiint(q[1])));
                        asm("min.f32 %0, %0, %1, %2;" : "+f"(accmin) : "f"(fx), "f"(fy));
                    }
                    t[u].x = q[0]; t[u].y = q[1];
                }
            }
        }
    }
    unsigned long long c1 = clock64(), g1 = gtimer();
    if (MODE == 3) ok = __float_as_uint(accmin) >= 0x03700000u;
    double s = ok ? 0.0 : 1.0;
#pragma unroll
    for (int u = 0; u < UN; ++u) s += t[u].x + t[u].y;
    out[blockIdx.x * THREADS + tid] = s;
    if (tid == 0) { stamps[2 * blockIdx.x] = c1 - c0; stamps[2 * blockIdx.x + 1] = g1 - g0; }
}

template <int MODE, int MINB>
void run(const char *name, int iters, const double *d_in, double *d_out, unsigned long long *d_st) {
    int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const int ctas_per_sm = MINB;
    const int grid = sms * ctas_per_sm;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    lab_kernel<MODE, MINB><<<grid, THREADS>>>(d_in, d_out, iters / 4, d_st);          // warm-up
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    lab_kernel<MODE, MINB><<<grid, THREADS>>>(d_in, d_out, iters, d_st);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms = 0; CK(cudaEventElapsedTime(&ms, e0, e1));
    std::vector<unsigned long long> st(2 * grid);
    CK(cudaMemcpy(st.data(), d_st, sizeof(unsigned long long) * 2 * grid, cudaMemcpyDeviceToHost));
    double cyc = 0, ns = 0;
    for (int i = 0; i < grid; ++i) { cyc += (double)st[2 * i]; ns += (double)st[2 * i + 1]; }
    const double mhz = cyc / ns * 1e3;
    const double cell_levels = (double)grid * THREADS * 16.0 * F * iters;
    const double fp64_warp_instr_per_smsp = cell_levels * 6.0 / 32.0 / (sms * 4.0);
    const double cycles = ms * 1e-3 * mhz * 1e6;
    printf("%-44s ctas/SM=%d  %8.3f ms  clock %6.0f MHz  fp64 pipe %5.1f %%  %7.2f Gcell-levels/s  (cfg4 pass of 8 levels: %6.3f ms)\n",
           name, ctas_per_sm, ms, mhz, 100.0 * fp64_warp_instr_per_smsp * 2.0 / cycles, cell_levels / ms * 1e-6,
           536920064.0 * 8.0 / (cell_levels / ms));
    fflush(stdout);
}

int main(int argc, char **argv) {
    const int iters = argc > 1 ? atoi(argv[1]) : 4000;
    std::vector<double> h(4096);
    srand(1);
    for (auto &v : h) v = 0.5 + (double)rand() / RAND_MAX;
    double *d_in, *d_out; unsigned long long *d_st;
    CK(cudaMalloc(&d_in, 4096 * 8)); CK(cudaMalloc(&d_out, 148 * 8 * THREADS * 8)); CK(cudaMalloc(&d_st, 148 * 8 * 16));
    CK(cudaMemcpy(d_in, h.data(), 4096 * 8, cudaMemcpyHostToDevice));
#define ALL(MB) \
    run<0, MB>("pure DFMA (6 per cell-level)", iters, d_in, d_out, d_st); \
    run<1, MB>("real mix, LDS operands, no guard", iters, d_in, d_out, d_st); \
    run<2, MB>("real mix + round-1 guard (LOP3 + 2 ISETP)", iters, d_in, d_out, d_st); \
    run<3, MB>("real mix + FMNMX3 guard (1 per 2 cells)", iters, d_in, d_out, d_st); \
    run<4, MB>("real mix + IMAD/ISETP guard (2 per cell)", iters, d_in, d_out, d_st);
    ALL(1) ALL(2) ALL(3) ALL(4)
    return 0;
}
