// dmf_rng.cuh — numpy's legacy generator on the device: the streams the reference draws per bootstrap
// resample, bit for bit.
//   sklearn.utils.resample(..., random_state=seed)  (bootstrap.py:28)   = RandomState(seed).randint(0, M, size=M)
//   set_seed(seed); rd.uniform(size=(M, n_u))       (deconvolution.py:9-11, :54-55)
// RandomState(seed) is MT19937 seeded by init_genrand (numpy/random/src/mt19937/mt19937.c: mt19937_seed); randint with a range
// below 2^32 is masked rejection on 32-bit outputs (numpy/random/src/distributions/distributions.c:
// buffered_bounded_masked_uint32 - "do val = next_uint32 & mask; while (val > rng)", no buffering at 32 bits); uniform is
// next_double = ((a >> 5) * 2^26 + (b >> 6)) / 2^53 of two consecutive outputs.  Third-party algorithm (numpy 1.26 / 2.x, identical
// in both), restated from its published source; pinned in tests/test_gpu_rng.py against numpy itself.
//
// One warp per (stream, task): the 624-word state lives in shared memory, a regeneration is 20 rounds of 32 lanes (element i needs
// old [i], old [i + 1] and [i + 397] (old, i < 227) or [i - 227] (already new)), the rejection step is a ballot + prefix count so the
// accepted values keep their order.
#pragma once
#include <cstdint>
#include "dmf_device.cuh"

namespace dmf {

constexpr int kMtN = 624, kMtM = 397;
constexpr int kRngWarps = 4;

__device__ __forceinline__ uint32_t mt_temper(uint32_t y) {
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
}

// next 624 words (mt19937_gen), whole warp
__device__ __forceinline__ void mt_regenerate(uint32_t* k, int lane) {
    for (int base = 0; base < kMtN - 1; base += 32) {
        const int i = base + lane;
        uint32_t v = 0;
        if (i < kMtN - 1) {
            const uint32_t y = (k[i] & 0x80000000u) | (k[i + 1] & 0x7fffffffu);
            const int m = i < kMtN - kMtM ? i + kMtM : i - (kMtN - kMtM);
            v = k[m] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        __syncwarp();
        if (i < kMtN - 1) k[i] = v;
        __syncwarp();
    }
    if (lane == 0) {
        const uint32_t y = (k[kMtN - 1] & 0x80000000u) | (k[0] & 0x7fffffffu);
        k[kMtN - 1] = k[kMtM - 1] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    __syncwarp();
}

// grid = (ceil(n_streams / kRngWarps), 2): blockIdx.y = 0 draws the resample indices, 1 the uniform doubles (both streams start from
// the SAME seed, as in the reference: resample() builds a fresh RandomState(seed) and init_BSSMF_md re-seeds the global one)
__global__ void __launch_bounds__(kRngWarps * 32) legacy_streams_kernel(const uint32_t* __restrict__ seeds, int n_streams, long long M,
                                                                       int32_t* __restrict__ idx, long long ld_idx, long long n_dbl,
                                                                       double* __restrict__ u, long long ld_u, uint32_t* __restrict__ state) {
    __shared__ uint32_t keys[kRngWarps][kMtN];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x * kRngWarps + warp;
    if (b >= n_streams) return;
    const int task = blockIdx.y;
    if (task == 0 && (idx == nullptr || M <= 0)) return;
    if (task == 1 && u == nullptr && state == nullptr) return;
    uint32_t* k = keys[warp];
    if (lane == 0) {                       // mt19937_seed
        uint32_t s = seeds[b];
        for (int p = 0; p < kMtN; ++p) {
            k[p] = s;
            s = 1812433253u * (s ^ (s >> 30)) + (uint32_t)p + 1u;
        }
    }
    __syncwarp();
    if (task == 0) {
        // randint(0, M): rng = M - 1, mask = smallest 2^k - 1 >= rng
        const uint32_t rng = (uint32_t)(M - 1);
        int32_t* out = idx + (size_t)b * ld_idx;
        if (rng == 0) {
            for (long long i = lane; i < M; i += 32) out[i] = 0;
            return;
        }
        uint32_t mask = rng;
        mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
        long long count = 0;
        while (count < M) {
            mt_regenerate(k, lane);
            for (int base = 0; base < kMtN && count < M; base += 32) {
                const int j = base + lane;
                const uint32_t v = j < kMtN ? (mt_temper(k[j]) & mask) : 0xffffffffu;
                const bool ok = j < kMtN && v <= rng;
                const unsigned bal = __ballot_sync(0xffffffffu, ok);
                const long long at = count + __popc(bal & ((1u << lane) - 1u));
                if (ok && at < M) out[at] = (int32_t)v;
                count += __popc(bal);
            }
            __syncwarp();
        }
    } else {
        double* out = u ? u + (size_t)b * ld_u : nullptr;
        long long done = 0;                // doubles produced so far
        int pos = kMtN;                    // numpy's state.pos after the last draw
        while (done < n_dbl) {
            mt_regenerate(k, lane);
            const long long left = n_dbl - done;
            const int take = (int)(left < kMtN / 2 ? left : kMtN / 2);
            if (out) {
                for (int p = lane; p < take; p += 32) {
                    const uint32_t a = mt_temper(k[2 * p]) >> 5, c = mt_temper(k[2 * p + 1]) >> 6;
                    out[done + p] = ((double)a * 67108864.0 + (double)c) / 9007199254740992.0;
                }
            }
            done += take;
            pos = 2 * take;
            __syncwarp();
        }
        if (state) {
            uint32_t* so = state + (size_t)b * (kMtN + 1);
            for (int p = lane; p < kMtN; p += 32) so[p] = k[p];
            if (lane == 0) so[kMtN] = (uint32_t)pos;
        }
    }
}

}  // namespace dmf
