// fp64_peak.cu — what the FP64 datapath of this GPU can issue, measured (the second roofline).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_peak tools/fp64_peak.cu && tools/fp64_peak [out.json]
// One JSON object per line:  {"test": ..., "warps_per_sm": W, "ilp": I, "tflops": T, "ops_per_clk_sm": O, "ms": t}
//   dfma          independent DFMA chains per thread (ILP) at several occupancies       -> FP64 FMA peak
//   dfma_dep      one dependent chain, one warp per SM sub-partition                      -> DFMA latency (clk)
//   dmma884       mma.sync.m8n8k4.f64 (DMMA), independent accumulators                   -> FP64 tensor peak
//   dmma16816     mma.sync.m16n8k16.f64 (sm_90+ shape)
//   mix           DFMA and DMMA interleaved in the same warps                             -> one pipe or two?
//   i2f           cvt.rn.f64.u32 chains, alone and interleaved with DFMA                  -> cost of the u16 coverage conversion
//   hmma / imma   legacy mma.sync bf16 m16n8k16 / s8 m16n8k32                             -> what an Ozaki-style split could use
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

template <int ILP>
__global__ void dfma_kernel(double* out, int iters, double a, double b) {
    double v[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) v[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int i = 0; i < ILP; ++i) v[i] = fma(v[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += v[i];
    if (s == 12345.678) out[0] = s;
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma16816(double (&c)[4], const double (&a)[8], const double (&b)[4]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

template <int ILP>
__global__ void dmma884_kernel(double* out, int iters, double a, double b) {
    double c[ILP][2];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int i = 0; i < ILP; ++i) dmma884(c[i][0], c[i][1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1];
    if (s == 12345.678) out[0] = s;
}

template <int ILP>
__global__ void dmma16816_kernel(double* out, int iters, double a0, double b0) {
    double c[ILP][4], a[8], b[4];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = a0 + i;
#pragma unroll
    for (int i = 0; i < 4; ++i) b[i] = b0 + i;
#pragma unroll
    for (int i = 0; i < ILP; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; c[i][2] = 1; c[i][3] = 2; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int i = 0; i < ILP; ++i) dmma16816(c[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    if (s == 12345.678) out[0] = s;
}

// NF DFMA per DMMA, both with independent accumulators
template <int NF, int ND>
__global__ void mix_kernel(double* out, int iters, double a, double b) {
    double v[NF > 0 ? NF : 1], c[ND > 0 ? ND : 1][2];
#pragma unroll
    for (int i = 0; i < NF; ++i) v[i] = threadIdx.x * 1e-3 + i;
#pragma unroll
    for (int i = 0; i < ND; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int i = 0; i < (NF > ND ? NF : ND); ++i) {
                if (i < ND) dmma884(c[i][0], c[i][1], a, b);
                if (i < NF) v[i] = fma(v[i], a, b);
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NF; ++i) s += v[i];
#pragma unroll
    for (int i = 0; i < ND; ++i) s += c[i][0] + c[i][1];
    if (s == 12345.678) out[0] = s;
}

// NI conversions u32 -> f64 and NF DFMA per round
template <int NI, int NF>
__global__ void i2f_kernel(double* out, int iters, double a, double b, unsigned seed) {
    unsigned u[NI];
    double v[NF > 0 ? NF : 1], acc[NI];
#pragma unroll
    for (int i = 0; i < NI; ++i) { u[i] = seed + threadIdx.x + i; acc[i] = 0; }
#pragma unroll
    for (int i = 0; i < NF; ++i) v[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int i = 0; i < NI; ++i) {
                double d;
                asm volatile("cvt.rn.f64.u32 %0, %1;" : "=d"(d) : "r"(u[i]));
                u[i] = __double2loint(d) ^ (unsigned)it;        // keeps the conversion inside the loop (one LOP + one MOV per cvt)
            }
#pragma unroll
            for (int i = 0; i < NF; ++i) v[i] = fma(v[i], a, b);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NI; ++i) s += (double)u[i] + acc[i];
#pragma unroll
    for (int i = 0; i < NF; ++i) s += v[i];
    if (s == 12345.678) out[0] = s;
}

__global__ void dfma_dep_kernel(double* out, long long* cycles, int iters, double a, double b) {
    double v = threadIdx.x;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 16; ++r) v = fma(v, a, b);
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
    if (v == 12345.678) out[0] = v;
}

template <int ILP>
__global__ void hmma_kernel(float* out, int iters, unsigned a0, unsigned b0) {
    float c[ILP][4];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; c[i][2] = 0; c[i][3] = 1; }
    unsigned a[4] = {a0, a0 + 1, a0 + 2, a0 + 3}, b[2] = {b0, b0 + 1};
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int i = 0; i < ILP; ++i)
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    if (s == 12345.678f) out[0] = s;
}
template <int ILP>
__global__ void imma_kernel(int* out, int iters, unsigned a0, unsigned b0) {
    int c[ILP][4];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; c[i][2] = 0; c[i][3] = 1; }
    unsigned a[4] = {a0, a0 + 1, a0 + 2, a0 + 3}, b[2] = {b0, b0 + 1};
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int i = 0; i < ILP; ++i)
                asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+r"(c[i][0]), "+r"(c[i][1]), "+r"(c[i][2]), "+r"(c[i][3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
    int s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    if (s == 123456789) out[0] = s;
}

static FILE* g_out = nullptr;
static int g_sms = 0;
static double g_mhz = 0;

template <typename F>
static double time_ms(F launch) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch(); launch();
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        CK(cudaEventRecord(e0));
        launch();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    return best;
}

static void report(const char* test, int warps_per_sm, int ilp, double ops /* MAC-equivalents (2 flop) or raw ops */, double ms, const char* unit) {
    const double per_s = ops / (ms * 1e-3);
    char line[512];
    snprintf(line, sizeof line, "{\"test\": \"%s\", \"warps_per_sm\": %d, \"ilp\": %d, \"ms\": %.4f, \"%s\": %.3f, \"ops_per_clk_sm_at_max_clock\": %.2f}", test,
             warps_per_sm, ilp, ms, unit, (unit[0] == 't' ? 2.0 : 1.0) * per_s / 1e12, per_s / (g_sms * g_mhz * 1e6));
    puts(line);
    if (g_out) { fputs(line, g_out); fputc('\n', g_out); fflush(g_out); }
}

int main(int argc, char** argv) {
    if (argc > 1) g_out = fopen(argv[1], "w");
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, 0));
    g_sms = p.multiProcessorCount;
    int khz = 0;
    CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0));
    g_mhz = khz / 1e3;
    printf("{\"device\": \"%s\", \"sms\": %d, \"max_sm_mhz\": %.0f}\n", p.name, g_sms, g_mhz);
    if (g_out) fprintf(g_out, "{\"device\": \"%s\", \"sms\": %d, \"max_sm_mhz\": %.0f}\n", p.name, g_sms, g_mhz);
    double* out; long long* cyc;
    CK(cudaMalloc(&out, 1024)); CK(cudaMalloc(&cyc, 64));
    const int iters = 4000;
    const int wlist[] = {4, 8, 16, 32};
    for (int w : wlist) {
        const int threads = 256, ctas = g_sms * (w * 32 / threads > 0 ? w * 32 / threads : 1), thr = w * 32 < threads ? w * 32 : threads;
        const double nthreads = (double)ctas * thr;
        report("dfma", w, 4, nthreads * iters * 8.0 * 4, time_ms([&] { dfma_kernel<4><<<ctas, thr>>>(out, iters, 1.0000001, 1e-9); }), "tflops");
        report("dfma", w, 8, nthreads * iters * 8.0 * 8, time_ms([&] { dfma_kernel<8><<<ctas, thr>>>(out, iters, 1.0000001, 1e-9); }), "tflops");
        report("dfma", w, 16, nthreads * iters * 8.0 * 16, time_ms([&] { dfma_kernel<16><<<ctas, thr>>>(out, iters, 1.0000001, 1e-9); }), "tflops");
        // one m8n8k4 = 256 MAC per warp = 8 per thread
        report("dmma884", w, 4, nthreads * iters * 4.0 * 4 * 8, time_ms([&] { dmma884_kernel<4><<<ctas, thr>>>(out, iters, 1.0000001, 1e-9); }), "tflops");
        report("dmma884", w, 8, nthreads * iters * 4.0 * 8 * 8, time_ms([&] { dmma884_kernel<8><<<ctas, thr>>>(out, iters, 1.0000001, 1e-9); }), "tflops");
        // one m16n8k16 = 2048 MAC per warp = 64 per thread
        report("dmma16816", w, 4, nthreads * iters * 2.0 * 4 * 64, time_ms([&] { dmma16816_kernel<4><<<ctas, thr>>>(out, iters, 1.0000001, 1e-9); }), "tflops");
        // mixes: MAC-equivalents = NF + 8 ND per round
        report("mix_8dfma_1dmma", w, 9, nthreads * iters * 4.0 * (8 + 8), time_ms([&] { mix_kernel<8, 1><<<ctas, thr>>>(out, iters, 1.0000001, 1e-9); }), "tflops");
        report("mix_8dfma_4dmma", w, 12, nthreads * iters * 4.0 * (8 + 32), time_ms([&] { mix_kernel<8, 4><<<ctas, thr>>>(out, iters, 1.0000001, 1e-9); }), "tflops");
        report("mix_4dfma_8dmma", w, 12, nthreads * iters * 4.0 * (4 + 64), time_ms([&] { mix_kernel<4, 8><<<ctas, thr>>>(out, iters, 1.0000001, 1e-9); }), "tflops");
        report("i2f_only", w, 8, nthreads * iters * 4.0 * 8, time_ms([&] { i2f_kernel<8, 0><<<ctas, thr>>>(out, iters, 1.0000001, 1e-9, 7u); }), "tera_cvt");
        report("i2f_2_plus_dfma_16", w, 18, nthreads * iters * 4.0 * 18, time_ms([&] { i2f_kernel<2, 16><<<ctas, thr>>>(out, iters, 1.0000001, 1e-9, 7u); }), "tera_ops");
        report("hmma_bf16_16816", w, 8, nthreads * iters * 4.0 * 8 * 128, time_ms([&] { hmma_kernel<8><<<ctas, thr>>>((float*)out, iters, 0x3f803f80u, 0x3f803f80u); }), "tflops");
        report("imma_s8_16832", w, 8, nthreads * iters * 4.0 * 8 * 256, time_ms([&] { imma_kernel<8><<<ctas, thr>>>((int*)out, iters, 0x01010101u, 0x01010101u); }), "tflops");
    }
    // dependent DFMA latency: one warp, cycles per DFMA
    dfma_dep_kernel<<<1, 32>>>(out, cyc, 2000, 1.0000001, 1e-9);
    CK(cudaDeviceSynchronize());
    long long hc;
    CK(cudaMemcpy(&hc, cyc, 8, cudaMemcpyDeviceToHost));
    printf("{\"test\": \"dfma_dependent_latency\", \"cycles_per_dfma\": %.2f}\n", (double)hc / (2000.0 * 16));
    if (g_out) { fprintf(g_out, "{\"test\": \"dfma_dependent_latency\", \"cycles_per_dfma\": %.2f}\n", (double)hc / (2000.0 * 16)); fclose(g_out); }
    return 0;
}
