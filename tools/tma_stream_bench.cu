// tma_stream_bench.cu — microbenchmark of the streaming skeleton used by the deconvolution kernels:
// how fast can a CTA-local ring of 1-D bulk copies (cp.async.bulk, SASS UBLKCP) pull a large array out of HBM,
// as a function of stage size, ring depth, copies per stage and CTAs per SM?  Also a plain LDG.128 reference.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_stream_bench tools/tma_stream_bench.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint32_t bar, uint32_t b) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(b) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t par) {
    uint32_t ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(bar), "r"(par) : "memory");
    return ok;
}
__device__ __forceinline__ void bulk(uint32_t dst, const void* src, uint32_t n, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(n), "r"(bar) : "memory");
}

// mode 0: dedicated producer warp (extra warp), mode 1: warp 0 produces between its own tiles
template <int MODE>
__global__ void ring_kernel(const char* __restrict__ src, size_t bytes, int stage_bytes, int stages, int split, int work, double* out) {
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned long long* full = reinterpret_cast<unsigned long long*>(smem);
    unsigned long long* empty = full + 16;
    unsigned char* buf = smem + 256;   // room for 16 + 16 barriers
    const int ncons = MODE == 0 ? blockDim.x - 32 : blockDim.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) { mbar_init(smem_u32(&full[s]), 1); mbar_init(smem_u32(&empty[s]), ncons / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const long long n_tiles = bytes / stage_bytes;
    const long long n_my = (n_tiles > blockIdx.x) ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    auto produce = [&](long long it) {
        if (it >= n_my) return;
        const int s = it % stages;
        const unsigned n = (unsigned)(it / stages);
        while (!mbar_try(smem_u32(&empty[s]), (n & 1u) ^ 1u)) {}
        if (lane == 0) {
            mbar_expect(smem_u32(&full[s]), stage_bytes);
            const char* g = src + (size_t)(blockIdx.x + it * gridDim.x) * stage_bytes;
            const int piece = stage_bytes / split;
            for (int k = 0; k < split; ++k) bulk(smem_u32(buf + (size_t)s * stage_bytes + k * piece), g + (size_t)k * piece, piece, smem_u32(&full[s]));
        }
        __syncwarp();
    };
    double acc = 0.0;
    if (MODE == 0 && warp == ncons / 32) {
        for (long long it = 0; it < n_my; ++it) produce(it);
    } else {
        if (MODE == 1 && warp == 0) for (int it = 0; it < stages - 1; ++it) produce(it);
        for (long long it = 0; it < n_my; ++it) {
            if (MODE == 1 && warp == 0) produce(it + stages - 1);
            const int s = it % stages;
            while (!mbar_try(smem_u32(&full[s]), (unsigned)(it / stages) & 1u)) {}
            const double* p = reinterpret_cast<const double*>(buf + (size_t)s * stage_bytes);
            // consume: every thread reads its share of the stage once; `work` extra dependent FMAs per element
            for (int i = threadIdx.x; i < stage_bytes / 8; i += ncons) {
                double v = p[i];
                for (int w = 0; w < work; ++w) v = fma(v, 1.0000001, 1e-9);
                acc += v;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&empty[s]));
        }
    }
    if (acc == 123.456) out[0] = acc;
}

__global__ void ldg_kernel(const double2* __restrict__ src, size_t n, int work, double* out) {
    double acc = 0.0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        double2 v = src[i];
        for (int w = 0; w < work; ++w) { v.x = fma(v.x, 1.0000001, 1e-9); v.y = fma(v.y, 1.0000001, 1e-9); }
        acc += v.x + v.y;
    }
    if (acc == 123.456) out[0] = acc;
}

template <int MODE>
float run(const char* d, size_t bytes, int stage_bytes, int stages, int split, int work, int ctas_per_sm, int threads, double* out) {
    if (stages > 16) { printf("stages > 16 unsupported\n"); exit(1); }
    const int smem = 256 + stage_bytes * stages;
    cudaFuncSetAttribute(ring_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e9;
    for (int r = 0; r < 4; ++r) {
        cudaEventRecord(a);
        ring_kernel<MODE><<<148 * ctas_per_sm, threads + (MODE == 0 ? 32 : 0), smem>>>(d, bytes, stage_bytes, stages, split, work, out);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (r && ms < best) best = ms;
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); exit(1); }
    return best;
}

int main() {
    const size_t bytes = (size_t)3 << 30;   // 3 GiB >> L2
    char* d; double* out;
    cudaMalloc(&d, bytes); cudaMalloc(&out, 8);
    cudaMemset(d, 0, bytes);
    printf("mode,stage_bytes,stages,split,ctas_per_sm,threads,work,ms,GBps\n");
    const int cfg[][6] = {  // stage_bytes, stages, split, ctas/SM, threads, work
        {8192, 4, 1, 2, 256, 0}, {16384, 4, 1, 2, 256, 0}, {21504, 4, 1, 2, 256, 0}, {21504, 4, 4, 2, 256, 0}, {24576, 4, 1, 2, 256, 0},
        {32768, 3, 1, 2, 256, 0}, {32768, 6, 1, 1, 256, 0}, {49152, 4, 1, 1, 256, 0}, {49152, 4, 1, 1, 512, 0}, {16384, 6, 1, 2, 256, 0},
        {8192, 8, 1, 2, 256, 0}, {8192, 12, 1, 2, 256, 0}, {4096, 16, 1, 2, 256, 0}, {16384, 4, 1, 3, 256, 0}, {8192, 6, 1, 4, 256, 0},
        {21504, 4, 1, 2, 256, 4}, {21504, 4, 1, 2, 256, 8}, {21504, 4, 1, 2, 256, 16}, {8192, 12, 1, 2, 256, 8}, {16384, 6, 1, 2, 256, 8},
    };
    for (auto& c : cfg) {
        for (int mode = 0; mode < 2; ++mode) {
            float ms = mode == 0 ? run<0>(d, bytes, c[0], c[1], c[2], c[5], c[3], c[4], out) : run<1>(d, bytes, c[0], c[1], c[2], c[5], c[3], c[4], out);
            const size_t used = bytes / c[0] * c[0];
            printf("%d,%d,%d,%d,%d,%d,%d,%.4f,%.1f\n", mode, c[0], c[1], c[2], c[3], c[4], c[5], ms, used / (ms * 1e-3) / 1e9);
        }
    }
    for (int work = 0; work <= 16; work += 8)
        for (int bpsm = 2; bpsm <= 8; bpsm *= 2) {
            cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
            float best = 1e9;
            for (int r = 0; r < 4; ++r) {
                cudaEventRecord(a);
                ldg_kernel<<<148 * bpsm, 256>>>((const double2*)d, bytes / 16, work, out);
                cudaEventRecord(b); cudaEventSynchronize(b);
                float ms; cudaEventElapsedTime(&ms, a, b);
                if (r && ms < best) best = ms;
            }
            printf("ldg,%d,0,0,%d,256,%d,%.4f,%.1f\n", 16, bpsm, work, best, bytes / (best * 1e-3) / 1e9);
        }
    return 0;
}
