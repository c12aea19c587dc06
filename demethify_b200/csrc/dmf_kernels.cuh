// dmf_kernels.cuh — reference-shaped streaming passes over the tall CpG dimension.
//
// Every kernel is ONE launch = one reference step for all live fits of a batch (grid.y = fit):
//   init_cost_kernel : cost_f_w (+ ||R||^2, max d_x at set-up)        deconvolution.py:15-17, 192-204, 218-221
//   u_pass_kernel    : one inner iteration of update_u                deconvolution.py:82-89  (unsup. variant :157-164)
//   alpha_pass_kernel: one inner iteration of update_alpha            deconvolution.py:94-101 + projection :21-37
//                      or one Frank-Wolfe iteration                   deconvolution.py:285-299
//
// Layout: a producer warp streams row tiles of X, D, R_trunc, u (and u_prev) into a 4-stage shared-memory
// ring with 1-D bulk copies (TMA engine) signalled through mbarriers; 256 consumer threads are arranged
// as (row group g, column thread tc): thread tc owns C adjacent sample columns, keeps its alpha columns in
// registers and walks the rows of the tile.  Row-wise sums (U gradient) use warp shuffles, column-wise
// sums (alpha gradient) stay in registers across the whole CTA lifetime; cross-CTA sums go through the
// deterministic two-level reduction of dmf_device.cuh and the LAST CTA applies the step (clip / simplex
// projection / Frank-Wolfe vertex), updates the fit state and re-arms the tickets.
#pragma once
#include "dmf_device.cuh"

namespace dmf {

// ------------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ T fma_t(T a, T b, T c);
template <>
__device__ __forceinline__ double fma_t<double>(double a, double b, double c) { return fma(a, b, c); }
template <>
__device__ __forceinline__ float fma_t<float>(float a, float b, float c) { return fmaf(a, b, c); }

// sum of one double per consumer thread, fixed order; result valid in consumer thread 0
__device__ __forceinline__ double consumer_block_sum(double v, double* scratch /* >= 8 doubles */, int ctid) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((ctid & 31) == 0) scratch[ctid >> 5] = v;
    consumer_bar();
    double s = 0.0;
    if (ctid == 0) {
#pragma unroll
        for (int w = 0; w < kConsumers / 32; ++w) s += scratch[w];
    }
    consumer_bar();
    return s;
}
__device__ __forceinline__ double consumer_block_max(double v, double* scratch, int ctid) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    if ((ctid & 31) == 0) scratch[ctid >> 5] = v;
    consumer_bar();
    double s = 0.0;
    if (ctid == 0) {
#pragma unroll
        for (int w = 0; w < kConsumers / 32; ++w) s = fmax(s, scratch[w]);
    }
    consumer_bar();
    return s;
}

// Common CTA set-up: barriers, tile sources.  Returns number of tiles this CTA owns.
struct CtaCtx {
    SmemCtl* ctl;
    char* stages;
    int n_my;
    int warp, lane, ctid;
};

__device__ __forceinline__ void cta_setup(const Geom& g, unsigned char* smem, CtaCtx& c) {
    c.ctl = reinterpret_cast<SmemCtl*>(smem);
    c.stages = reinterpret_cast<char*>(smem) + kCtlBytes;
    c.warp = threadIdx.x >> 5;
    c.lane = threadIdx.x & 31;
    c.ctid = threadIdx.x;   // consumers are threads [0, kConsumers)
    const int first = blockIdx.x;
    c.n_my = (g.n_tiles > first) ? (g.n_tiles - first + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < kStages; ++s) {
            mbar_init(smem_u32(&c.ctl->full[s]), 1);
            mbar_init(smem_u32(&c.ctl->empty[s]), kConsumers / 32);
        }
        mbar_fence_init();
    }
    __syncthreads();
}

__device__ __forceinline__ void producer_loop(const Geom& g, const FitDev& f, const CtaCtx& c, const TileSrc* src, int nsrc) {
    for (int it = 0; it < c.n_my; ++it) {
        const int s = it % kStages;
        const unsigned n = (unsigned)(it / kStages);
        mbar_wait(smem_u32(&c.ctl->empty[s]), (n & 1u) ^ 1u);
        const long long tile = blockIdx.x + (long long)it * gridDim.x;
        const long long r0 = tile * g.tile_rows;
        const int nrows = (int)min((long long)g.tile_rows, g.M - r0);
        produce_tile(src, nsrc, f.rows, r0, nrows, c.stages + (size_t)s * g.stage_bytes, smem_u32(&c.ctl->full[s]), c.lane);
    }
}

// ------------------------------------------------------------------------------------------------
// cost / set-up pass
template <typename T, typename WT, int KTB, int C>
__global__ void __launch_bounds__(kThreads, (KTB * C <= 16) ? 2 : 1) init_cost_kernel(const PassArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    const Geom& g = a.g;
    const FitDev f = a.fits[blockIdx.y];
    FitState* st = f.st;
    if (st->done) return;
    CtaCtx c;
    cta_setup(g, smem, c);
    const bool initial = (a.flags & kFlagInitial) != 0;
    const int ucur = st->u_cur, acur = st->a_cur;
    const T* Acur = reinterpret_cast<const T*>(f.A) + (size_t)acur * g.Kt * g.N;
    const char* Ucur = f.U + (size_t)ucur * g.uslot_bytes;

    double cost = 0.0, ssq_r = 0.0, ssq_u = 0.0, dmx = 0.0;
    if (c.warp == kConsumers / 32) {
        TileSrc src[4];
        src[0] = {f.X, g.ldx * (long long)sizeof(T), g.offX, true, (g.row_bulk & 1u) != 0};
        src[1] = {f.D, g.ldd * (long long)sizeof(WT), g.offD, true, (g.row_bulk & 2u) != 0};
        src[2] = {g.K ? f.Rk : nullptr, g.ldr * (long long)sizeof(T), g.offR, true, (g.row_bulk & 4u) != 0};
        src[3] = {Ucur, g.nu * (long long)sizeof(T), g.offU, false, false};
        producer_loop(g, f, c, src, 4);
    } else {
        const int tc = c.ctid % g.ntc, gr = c.ctid / g.ntc;
        const int j0 = tc * C;
        T at[KTB][C];
#pragma unroll
        for (int k = 0; k < KTB; ++k)
#pragma unroll
            for (int cc = 0; cc < C; ++cc) at[k][cc] = (k < g.Kt && j0 + cc < g.N) ? Acur[(size_t)k * g.N + j0 + cc] : (T)0;
        for (int it = 0; it < c.n_my; ++it) {
            const int s = it % kStages;
            mbar_wait(smem_u32(&c.ctl->full[s]), (unsigned)(it / kStages) & 1u);
            const char* sb = c.stages + (size_t)s * g.stage_bytes;
            const T* sX = reinterpret_cast<const T*>(sb + g.offX);
            const void* sD = sb + g.offD;
            const T* sR = reinterpret_cast<const T*>(sb + g.offR);
            const T* sU = reinterpret_cast<const T*>(sb + g.offU);
            const long long r0 = (blockIdx.x + (long long)it * gridDim.x) * g.tile_rows;
            const int nrows = (int)min((long long)g.tile_rows, g.M - r0);
            for (int r = gr; r < nrows; r += g.rg) {
                T rrow[KTB];
#pragma unroll
                for (int k = 0; k < KTB; ++k) {
                    T v = (T)0;
                    if (k < g.K) v = sR[(size_t)r * g.ldr + k];
                    else if (k < g.Kt) v = sU[r * g.nu + (k - g.K)];
                    rrow[k] = v;
                }
                if (initial && tc == 0) {
#pragma unroll
                    for (int k = 0; k < KTB; ++k) {
                        const double v = (double)rrow[k];
                        if (k < g.K) ssq_r = fma(v, v, ssq_r);
                        else ssq_u = fma(v, v, ssq_u);
                    }
                }
#pragma unroll
                for (int cc = 0; cc < C; ++cc) {
                    if (j0 + cc < g.N) {
                        const T x = sX[(size_t)r * g.ldx + j0 + cc];
                        const T d = wload<T, WT>(sD, (long long)r * g.ldd + j0 + cc);
                        T pred = (T)0;
#pragma unroll
                        for (int k = 0; k < KTB; ++k) pred = fma_t<T>(rrow[k], at[k][cc], pred);
                        const double res = (double)(x - pred);
                        cost = fma((double)d * res, res, cost);
                        dmx = fmax(dmx, (double)d);
                    }
                }
            }
            __syncwarp();
            if (c.lane == 0) mbar_arrive(smem_u32(&c.ctl->empty[s]));
        }
    }
    __syncthreads();
    // CTA partial record: [cost, ssq_rk, ssq_u, dmax]
    double* scratch = reinterpret_cast<double*>(c.stages);
    double* rec = scratch + 16;
    if (c.warp < kConsumers / 32) {
        const double s0 = consumer_block_sum(cost, scratch, c.ctid);
        const double s1 = consumer_block_sum(ssq_r, scratch, c.ctid);
        const double s2 = consumer_block_sum(ssq_u, scratch, c.ctid);
        const double s3 = consumer_block_max(dmx, scratch, c.ctid);
        if (c.ctid == 0) {
            double* p = f.part + (size_t)blockIdx.x * g.part_stride;
            p[0] = s0; p[1] = s1; p[2] = s2; p[3] = s3;
        }
    }
    // dmax needs a max, not a sum: it travels through slot 3 of every record and is re-derived below.
    if (!hier_reduce(g, f, rec, 3, &c.ctl->flag)) return;
    if (threadIdx.x == 0) {
        const double cf = rec[0];
        if (initial) {
            double dm = 0.0;
            for (int p = 0; p < g.n_parts; ++p) dm = fmax(dm, __ldcg(&f.part[(size_t)p * g.part_stride + 3]));
            st->dmax = dm;
            st->dmax2 = dm * dm;
            st->ssq_rk = rec[1];
            st->ssq_u = rec[2];
            // ||alpha[-n_u:]||_F^2 (deconvolution.py:198): tiny, done serially in row-major order
            double sa = 0.0;
            for (int q = 0; q < g.nu; ++q)
                for (int j = 0; j < g.N; ++j) {
                    const double v = (double)Acur[(size_t)(g.K + q) * g.N + j];
                    sa = fma(v, v, sa);
                }
            const double na = sqrt(sa), nr = sqrt(rec[1] + rec[2]);
            st->l_w = (na * na) * st->dmax2;
            st->l_w_old = st->l_w;
            st->l_h = (nr * nr) * st->dmax2;
            st->l_h_old = st->l_h;
            st->a1 = 1.0;
            st->a2 = 1.0;
            st->cf = cf;
            st->cf_prev = cf;
            st->n_outer = 0;
            if (f.trace && f.trace_cap > 0) f.trace[0] = cf;
        } else {
            const double prev = st->cf;
            st->cf_prev = prev;
            st->cf = cf;
            const int n = st->n_outer + 1;
            st->n_outer = n;
            if (f.trace && n < f.trace_cap) f.trace[n] = cf;
            if (fabs(cf - prev) < a.tol) st->done = 1;      // deconvolution.py:220
            if (!(cf == cf)) st->done = 3;                    // NaN guard
        }
    }
}

// ------------------------------------------------------------------------------------------------
// U pass: u <- clip(u_t + ((d o (x - Rk a_k - u_g a_u)) a_u^T) / l_w, 0, 1),  u_t = u + beta (u - u_prev)
// (u_g = u_t in update_u:88, u_g = u in unsupervised_deconv:163)
template <typename T, typename WT, int KTB, int NUB, int C>
__global__ void __launch_bounds__(kThreads, (KTB * C <= 16) ? 2 : 1) u_pass_kernel(const PassArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    const Geom& g = a.g;
    const FitDev f = a.fits[blockIdx.y];
    FitState* st = f.st;
    if (st->done) return;
    CtaCtx c;
    cta_setup(g, smem, c);
    const int ucur = st->u_cur, acur = st->a_cur;
    const double a_prev = st->a1, l_w = st->l_w, l_w_old = st->l_w_old;
    const double a_next = next_momentum(a_prev);
    const T beta = (T)extrap_beta(a_prev, a_next, l_w_old, l_w);
    const T lw = (T)l_w;
    const T* Acur = reinterpret_cast<const T*>(f.A) + (size_t)acur * g.Kt * g.N;
    const size_t uslot = (size_t)g.uslot_bytes;
    const char* Ucur = f.U + (size_t)ucur * uslot;
    char* Uprev = f.U + (size_t)(ucur ^ 1) * uslot;      // read as u_prev, overwritten with the new u
    const bool at_current = (g.mode == 2);

    double ssq_u = 0.0;
    if (c.warp == kConsumers / 32) {
        TileSrc src[5];
        src[0] = {f.X, g.ldx * (long long)sizeof(T), g.offX, true, (g.row_bulk & 1u) != 0};
        src[1] = {f.D, g.ldd * (long long)sizeof(WT), g.offD, true, (g.row_bulk & 2u) != 0};
        src[2] = {g.K ? f.Rk : nullptr, g.ldr * (long long)sizeof(T), g.offR, true, (g.row_bulk & 4u) != 0};
        src[3] = {Ucur, g.nu * (long long)sizeof(T), g.offU, false, false};
        src[4] = {Uprev, g.nu * (long long)sizeof(T), g.offUp, false, false};
        producer_loop(g, f, c, src, 5);
    } else {
        const int tc = c.ctid % g.ntc, gr = c.ctid / g.ntc;
        const int j0 = tc * C;
        const int lanes_per_row = min(g.ntc, 32);
        const int warps_per_row = (g.ntc + 31) / 32;
        T at[KTB][C];    // all alpha rows of my columns (prediction)
        T au[NUB][C];    // unknown block again (gradient), compile-time indexed
#pragma unroll
        for (int k = 0; k < KTB; ++k)
#pragma unroll
            for (int cc = 0; cc < C; ++cc) at[k][cc] = (k < g.Kt && j0 + cc < g.N) ? Acur[(size_t)k * g.N + j0 + cc] : (T)0;
#pragma unroll
        for (int q = 0; q < NUB; ++q)
#pragma unroll
            for (int cc = 0; cc < C; ++cc) au[q][cc] = (q < g.nu && j0 + cc < g.N) ? Acur[(size_t)(g.K + q) * g.N + j0 + cc] : (T)0;
        // cross-warp combine buffer (only when a row spans several warps): [2][tile_rows][warps_per_row][nu]
        double* red = reinterpret_cast<double*>(c.stages + (size_t)kStages * g.stage_bytes);
        const int rows_padded = ((g.tile_rows + g.rg - 1) / g.rg) * g.rg;

        for (int it = 0; it < c.n_my; ++it) {
            const int s = it % kStages;
            mbar_wait(smem_u32(&c.ctl->full[s]), (unsigned)(it / kStages) & 1u);
            const char* sb = c.stages + (size_t)s * g.stage_bytes;
            const T* sX = reinterpret_cast<const T*>(sb + g.offX);
            const void* sD = sb + g.offD;
            const T* sR = reinterpret_cast<const T*>(sb + g.offR);
            const T* sU = reinterpret_cast<const T*>(sb + g.offU);
            const T* sUp = reinterpret_cast<const T*>(sb + g.offUp);
            const long long r0 = (blockIdx.x + (long long)it * gridDim.x) * g.tile_rows;
            const int nrows = (int)min((long long)g.tile_rows, g.M - r0);
            double* redt = red + (size_t)(it & 1) * g.tile_rows * warps_per_row * g.nu;
            T* Uout = reinterpret_cast<T*>(Uprev) + (size_t)r0 * g.nu;

            for (int rb = 0; rb < rows_padded; rb += g.rg) {
                const int r = rb + gr;
                const bool live = r < nrows;
                T rrow[KTB];     // [Rk row | u_g row]
                T ut[NUB];
#pragma unroll
                for (int q = 0; q < NUB; ++q) {
                    T v = (T)0;
                    if (live && q < g.nu) {
                        const T u = sU[r * g.nu + q], up = sUp[r * g.nu + q];
                        v = u + beta * (u - up);
                    }
                    ut[q] = v;
                }
#pragma unroll
                for (int k = 0; k < KTB; ++k) {
                    T v = (T)0;
                    if (live) {
                        if (k < g.K) v = sR[(size_t)r * g.ldr + k];
                        else if (k < g.Kt) v = at_current ? sU[r * g.nu + (k - g.K)] : (T)0;
                    }
                    rrow[k] = v;
                }
                T gp[NUB];
#pragma unroll
                for (int q = 0; q < NUB; ++q) gp[q] = (T)0;
#pragma unroll
                for (int cc = 0; cc < C; ++cc) {
                    if (live && j0 + cc < g.N) {
                        const T x = sX[(size_t)r * g.ldx + j0 + cc];
                        const T d = wload<T, WT>(sD, (long long)r * g.ldd + j0 + cc);
                        T pk = (T)0;     // R_trunc @ alpha_known   (and u @ alpha_unk in the unsupervised variant)
#pragma unroll
                        for (int k = 0; k < KTB; ++k) pk = fma_t<T>(rrow[k], at[k][cc], pk);
                        T res = x - pk;
                        if (!at_current) {
                            T pu = (T)0;  // u_temp @ alpha_unk
#pragma unroll
                            for (int q = 0; q < NUB; ++q) pu = fma_t<T>(ut[q], au[q][cc], pu);
                            res = res - pu;
                        }
                        const T w = d * res;
#pragma unroll
                        for (int q = 0; q < NUB; ++q) gp[q] = fma_t<T>(w, au[q][cc], gp[q]);
                    }
                }
                // row sum over the column threads of this row
#pragma unroll
                for (int q = 0; q < NUB; ++q) {
                    if (q < g.nu) {
                        double v = (double)gp[q];
                        for (int o = lanes_per_row >> 1; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                        gp[q] = (T)v;
                        if (warps_per_row > 1 && live && (c.lane == 0)) redt[((size_t)r * warps_per_row + (tc >> 5)) * g.nu + q] = v;
                    }
                }
                if (warps_per_row == 1 && live && (tc == 0)) {
#pragma unroll
                    for (int q = 0; q < NUB; ++q) {
                        if (q < g.nu) {
                            T un = ut[q] + gp[q] / lw;
                            un = un < (T)0 ? (T)0 : (un > (T)1 ? (T)1 : un);
                            Uout[(size_t)r * g.nu + q] = un;
                            ssq_u = fma((double)un, (double)un, ssq_u);
                        }
                    }
                }
            }
            if (warps_per_row > 1) {
                consumer_bar();
                for (int e = c.ctid; e < nrows * g.nu; e += kConsumers) {
                    const int r = e / g.nu, q = e - r * g.nu;
                    double v = 0.0;
                    for (int w = 0; w < warps_per_row; ++w) v += redt[((size_t)r * warps_per_row + w) * g.nu + q];
                    const T u = sU[r * g.nu + q], up = sUp[r * g.nu + q];
                    const T utq = u + beta * (u - up);
                    T un = utq + (T)v / lw;
                    un = un < (T)0 ? (T)0 : (un > (T)1 ? (T)1 : un);
                    Uout[(size_t)r * g.nu + q] = un;
                    ssq_u = fma((double)un, (double)un, ssq_u);
                }
            }
            __syncwarp();
            if (c.lane == 0) mbar_arrive(smem_u32(&c.ctl->empty[s]));
        }
    }
    __syncthreads();
    double* scratch = reinterpret_cast<double*>(c.stages);
    double* rec = scratch + 16;
    if (c.warp < kConsumers / 32) {
        const double s0 = consumer_block_sum(ssq_u, scratch, c.ctid);
        if (c.ctid == 0) f.part[(size_t)blockIdx.x * g.part_stride] = s0;
    }
    if (!hier_reduce(g, f, rec, 1, &c.ctl->flag)) return;
    if (threadIdx.x == 0) {
        st->a1 = a_next;
        st->l_w_old = l_w;                                   // deconvolution.py:89
        st->u_cur = ucur ^ 1;
        st->ssq_u = rec[0];
        const double nr = sqrt(st->ssq_rk + rec[0]);
        st->l_h = (nr * nr) * st->dmax2;                     // deconvolution.py:212 (read by the alpha steps only)
    }
}

// ------------------------------------------------------------------------------------------------
// simplex projection of one column held in v[0..p) (deconvolution.py:21-37); returns false on NaN
__device__ __forceinline__ bool project_simplex(double* v, int p) {
    double u[kMaxKt];
    for (int i = 0; i < p; ++i) u[i] = v[i];
    for (int i = 1; i < p; ++i) {          // insertion sort, descending
        const double key = u[i];
        int j = i - 1;
        while (j >= 0 && u[j] < key) { u[j + 1] = u[j]; --j; }
        u[j + 1] = key;
    }
    double cs = 0.0, theta = 0.0;
    int rho = -1;
    for (int j = 0; j < p; ++j) {
        cs += u[j];
        const double pi = cs - 1.0;
        if (u[j] - pi / (double)(j + 1) > 0.0) { rho = j; theta = pi / (double)(j + 1); }
    }
    if (rho < 0) return false;
    for (int i = 0; i < p; ++i) v[i] = fmax(v[i] - theta, 0.0);
    return true;
}

// alpha pass: G = R^T (d o (x - R a_eval)); last CTA applies the projected-gradient or Frank-Wolfe step
template <typename T, typename WT, int KTB, int C>
__global__ void __launch_bounds__(kThreads, (KTB * C <= 16) ? 2 : 1) alpha_pass_kernel(const PassArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    const Geom& g = a.g;
    const FitDev f = a.fits[blockIdx.y];
    FitState* st = f.st;
    if (st->done) return;
    CtaCtx c;
    cta_setup(g, smem, c);
    const bool fw = (a.flags & kFlagFW) != 0;
    const int ucur = st->u_cur, acur = st->a_cur;
    const double a_prev = st->a2, l_h = st->l_h, l_h_old = st->l_h_old;
    const double a_next = next_momentum(a_prev);
    const double beta_d = fw ? 0.0 : extrap_beta(a_prev, a_next, l_h_old, l_h);
    const T beta = (T)beta_d;
    T* Acur = reinterpret_cast<T*>(f.A) + (size_t)acur * g.Kt * g.N;
    T* Aprev = reinterpret_cast<T*>(f.A) + (size_t)(acur ^ 1) * g.Kt * g.N;
    const char* Ucur = f.U + (size_t)ucur * g.uslot_bytes;

    const int tc = c.ctid % g.ntc, gr = c.ctid / g.ntc;
    const int j0 = tc * C;
    T G[KTB][C];
#pragma unroll
    for (int k = 0; k < KTB; ++k)
#pragma unroll
        for (int cc = 0; cc < C; ++cc) G[k][cc] = (T)0;

    if (c.warp == kConsumers / 32) {
        TileSrc src[4];
        src[0] = {f.X, g.ldx * (long long)sizeof(T), g.offX, true, (g.row_bulk & 1u) != 0};
        src[1] = {f.D, g.ldd * (long long)sizeof(WT), g.offD, true, (g.row_bulk & 2u) != 0};
        src[2] = {g.K ? f.Rk : nullptr, g.ldr * (long long)sizeof(T), g.offR, true, (g.row_bulk & 4u) != 0};
        src[3] = {Ucur, g.nu * (long long)sizeof(T), g.offU, false, false};
        producer_loop(g, f, c, src, 4);
    } else {
        T at[KTB][C];    // evaluation point: alpha_temp (PG) or alpha (FW)
#pragma unroll
        for (int k = 0; k < KTB; ++k)
#pragma unroll
            for (int cc = 0; cc < C; ++cc) {
                T v = (T)0;
                if (k < g.Kt && j0 + cc < g.N) {
                    const T ac = Acur[(size_t)k * g.N + j0 + cc];
                    v = ac;
                    if (!fw) v = ac + beta * (ac - Aprev[(size_t)k * g.N + j0 + cc]);
                }
                at[k][cc] = v;
            }
        for (int it = 0; it < c.n_my; ++it) {
            const int s = it % kStages;
            mbar_wait(smem_u32(&c.ctl->full[s]), (unsigned)(it / kStages) & 1u);
            const char* sb = c.stages + (size_t)s * g.stage_bytes;
            const T* sX = reinterpret_cast<const T*>(sb + g.offX);
            const void* sD = sb + g.offD;
            const T* sR = reinterpret_cast<const T*>(sb + g.offR);
            const T* sU = reinterpret_cast<const T*>(sb + g.offU);
            const long long r0 = (blockIdx.x + (long long)it * gridDim.x) * g.tile_rows;
            const int nrows = (int)min((long long)g.tile_rows, g.M - r0);
            for (int r = gr; r < nrows; r += g.rg) {
                T rrow[KTB];
#pragma unroll
                for (int k = 0; k < KTB; ++k) {
                    T v = (T)0;
                    if (k < g.K) v = sR[(size_t)r * g.ldr + k];
                    else if (k < g.Kt) v = sU[r * g.nu + (k - g.K)];
                    rrow[k] = v;
                }
#pragma unroll
                for (int cc = 0; cc < C; ++cc) {
                    if (j0 + cc < g.N) {
                        const T x = sX[(size_t)r * g.ldx + j0 + cc];
                        const T d = wload<T, WT>(sD, (long long)r * g.ldd + j0 + cc);
                        T pred = (T)0;
#pragma unroll
                        for (int k = 0; k < KTB; ++k) pred = fma_t<T>(rrow[k], at[k][cc], pred);
                        const T w = d * (x - pred);
#pragma unroll
                        for (int k = 0; k < KTB; ++k) G[k][cc] = fma_t<T>(rrow[k], w, G[k][cc]);
                    }
                }
            }
            __syncwarp();
            if (c.lane == 0) mbar_arrive(smem_u32(&c.ctl->empty[s]));
        }
    }
    __syncthreads();
    // combine the row groups of this CTA in fixed order, then publish the CTA partial [Kt][N]
    double* scratch = reinterpret_cast<double*>(c.stages);
    const int KN = g.Kt * g.N;
    if (c.warp < kConsumers / 32) {
        for (int gg = 0; gg < g.rg; ++gg) {
            if (gr == gg) {
#pragma unroll
                for (int k = 0; k < KTB; ++k)
#pragma unroll
                    for (int cc = 0; cc < C; ++cc)
                        if (k < g.Kt && j0 + cc < g.N) {
                            double* p = &scratch[(size_t)k * g.N + j0 + cc];
                            *p = (gg == 0) ? (double)G[k][cc] : (*p + (double)G[k][cc]);
                        }
            }
            consumer_bar();
        }
        double* part = f.part + (size_t)blockIdx.x * g.part_stride;
        for (int e = c.ctid; e < KN; e += kConsumers) part[e] = scratch[e];
    }
    if (!hier_reduce(g, f, scratch, KN, &c.ctl->flag)) return;

    // ---- last CTA: apply the step on the Kt x N gradient held in scratch
    double* colred = scratch + KN;        // per-column ||alpha_unk||^2 contributions
    int bad = 0;
    for (int j = threadIdx.x; j < g.N; j += blockDim.x) {
        double v[kMaxKt];
        double sa = 0.0;
        if (!fw) {
            for (int k = 0; k < g.Kt; ++k) {
                const T ac = Acur[(size_t)k * g.N + j], ap = Aprev[(size_t)k * g.N + j];
                const T atv = ac + beta * (ac - ap);
                v[k] = (double)(atv + (T)(scratch[(size_t)k * g.N + j]) / (T)l_h);
            }
            if (!project_simplex(v, g.Kt)) bad = 1;
            else
                for (int k = 0; k < g.Kt; ++k) Aprev[(size_t)k * g.N + j] = (T)v[k];   // becomes the current slot
            for (int q = 0; q < g.nu; ++q) sa = fma(v[g.K + q], v[g.K + q], sa);
        } else {
            // gradient = -G; vertex = FIRST argmin of each block (np.argmin), deconvolution.py:286-299
            const double pj = f.purity[j];
            const double gamma = 2.0 / (double)(a.k_inner + 2);
            int i1 = 0, i2 = 0;
            double m1 = 0.0, m2 = 0.0;
            for (int k = 0; k < g.K; ++k) {
                const double gk = -scratch[(size_t)k * g.N + j];
                if (k == 0 || gk < m1) { m1 = gk; i1 = k; }
            }
            for (int q = 0; q < g.nu; ++q) {
                const double gq = -scratch[(size_t)(g.K + q) * g.N + j];
                if (q == 0 || gq < m2) { m2 = gq; i2 = q; }
            }
            for (int k = 0; k < g.Kt; ++k) {
                const double s = (k < g.K) ? ((k == i1) ? pj : 0.0) : ((k - g.K == i2) ? (1.0 - pj) : 0.0);
                const double an = (1.0 - gamma) * (double)Acur[(size_t)k * g.N + j] + gamma * s;
                Acur[(size_t)k * g.N + j] = (T)an;
                if (k >= g.K) sa = fma(an, an, sa);
            }
        }
        colred[j] = sa;
    }
    if (bad) st->done = 3;
    __syncthreads();
    if (threadIdx.x == 0) {
        double sa = 0.0;
        for (int j = 0; j < g.N; ++j) sa += colred[j];
        const double na = sqrt(sa);
        st->l_w = (na * na) * st->dmax2;                 // deconvolution.py:216 / :327 (read by the U steps only)
        if (!fw) {
            st->a2 = a_next;
            st->l_h_old = l_h;                           // deconvolution.py:101
            st->a_cur = acur ^ 1;
        }
    }
}

}  // namespace dmf
