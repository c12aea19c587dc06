// dmf_fused.cuh — "fused" engine: ONE streaming pass over X, d_x, R_trunc, u per OUTER iteration of
// deconvolution.py:206-221 (SURVEY.md 7.1 #1-#3, 8 d4 "fused outer iteration").
//
// The Gram-form engine (dmf_gram.cuh) needs two passes per outer iteration because the panel statistics of the alpha step
// need the u that the U step produces.  The U step is row-local, so a row tile can go through all three stages on one visit:
//
//   A  row statistics   c = x - R_trunc a_k,  b_m = (d o c) a_u^T,  H_m = a_u diag(d_m) a_u^T   (+ sum d c^2 for the cost)
//   U  n_iter2 update_u iterations per row on (b_m, H_m)  (deconvolution.py:82-89 / :157-164), the row's cost cross terms
//      -2 u^T b_m + u^T H_m u of the INCOMING iterate, the new (u, u_) written to the other slot pair
//   C  Gram panel with the NEW u:  G_uk += d (u (x) R_trunc),  G_uu += d (u (x) u),  bx_u += (d o x) u
//
// The cost of the incoming iterate (u_k, alpha_k) is complete at the end of the pass; the last CTA then runs the termination
// test of deconvolution.py:218-221 for outer iteration k and COMMITS the U step (flips the slot pair, publishes the panel,
// updates l_h, a1) only if the fit did not just terminate - so a terminated fit returns exactly the iterate whose cost passed
// the test.  alpha_inner_kernel (dmf_gram.cuh) follows and produces alpha_{k+1}.
//
// One CTA = 20 warps with four roles (warp specialised, tiles flow A -> U -> C through a 5-stage shared-memory ring):
//   8 A-warps   each owns 8 S samples; c comes from FP64 tensor-core MMAs (mma.sync.m8n8k4.f64, SASS DMMA) with the samples on
//               the M side, so a lane owns ONE sample and two rows per 8 x 8 block and keeps only its own alpha values
//   3 U-warps   lane = row, tile t goes to U-warp t mod 3 (the 20 dependent iterations of a tile take about as long as the
//               A and C stages of a tile together, so consecutive tiles must overlap)
//   1 producer  drives the TMA ring: per-row bulk copies into bank-conflict-free padded rows
//   8 C-warps   the panel is the GEMM  [N x rows] (d) x [rows x NCOL] (u (x) [R_trunc | u]) : DMMA again, accumulators stay in
//               the MMA C fragments for the whole kernel; bx_u on the FMA pipe
// FP64 only (tcgen05 has no FP64 kind; DMMA and DFMA share one pipe at 64 FMA/clk/SM - tools/fp64_peak.cu), which is the
// binding roofline of this kernel: ~39 FMA per (row, sample) against 10 bytes.
#pragma once
#include <type_traits>
#include "dmf_gram.cuh"

namespace dmf {

constexpr int kFA = 8, kFC = 8, kFU = 3, kFP = 1;
constexpr int kFusedThreads = (kFA + kFC + kFU + kFP) * 32;
constexpr int kFusedRows = 16;           // rows per tile (two 8-row MMA blocks, four 4-row k-steps)
constexpr int kFusedStages = 5;
constexpr int kFusedMaxInner = 64;       // beyond this the U-warps (16 rows at a time) would bound the pass: Gram engine instead
constexpr unsigned kFusedCtlBytes = 2048;

struct FusedArgs {
    Geom g;                  // problem sizes, pitches, n_parts / n_groups / part_stride of this launch
    const FitDev* fits;
    int n_iter2, flags;
    double tol;
    const double* mom_a;
    const double* mom_m;
    int n_tiles;
    unsigned pitchX, pitchD;                          // bytes per row of X / d_x, in global and in shared memory
    unsigned offD, offR, offU, offUp, stage_bytes;    // stage layout, X at 0
    unsigned offStats;                                // from the start of dynamic shared memory
    unsigned zero_off;                                // offset (inside a stage) of 8 bytes that always hold 0.0
};
typedef void (*fused_kern_t)(const FusedArgs);

struct FusedCtl {
    unsigned long long full[8], empty[8], stats[8], udone[8], sfree[2];
    double wsum[2][kFA + kFC + kFU + kFP];
    double beta[kFusedMaxInner];          // extrapolation weights of this launch's update_u iterations (the same for every row)
    int flag, commit;
};
static_assert(sizeof(FusedCtl) <= kFusedCtlBytes, "fused control block too large");

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <typename WT>
__device__ __forceinline__ double lds_weight(uint32_t addr);
template <>
__device__ __forceinline__ double lds_weight<uint16_t>(uint32_t addr) {
    uint16_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr));
    return (double)(unsigned)v;
}
template <>
__device__ __forceinline__ double lds_weight<double>(uint32_t addr) {
    double v;
    lds1(addr, v);
    return v;
}

// Row sums over the 8 lanes that share (lane & 3): halving butterfly over lane bits 2, 3, 4.  NV is a multiple of 8; on return
// the lane holds the totals of slots base .. base + NV / 8 - 1 in v[0 .. NV / 8).  (The row <-> MMA column mapping is a property
// of the whole warp, so the slots cannot be made lane dependent to save the selects.)
template <int NV>
__device__ __forceinline__ int reduce_over_groups(double (&v)[NV], int lane) {
    static_assert(NV % 8 == 0, "NV must be a multiple of 8");
    int base = 0;
#pragma unroll
    for (int s = 0; s < 3; ++s) {
        const int o = 4 << s;
        const int half = NV >> (s + 1);
        const bool up = (lane & o) != 0;
#pragma unroll
        for (int i = 0; i < half; ++i) {
            const double send = up ? v[i] : v[i + half];
            const double keep = up ? v[i + half] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
        }
        base += up ? half : 0;
    }
    return base;
}

// One update_u step of deconvolution.py:82-89 (gradient at u for the unsupervised variant, :163) on the row's statistics,
// (prev, cur) -> next.  The U-warps run 16 rows x n_iter2 of these back to back while the other 16 warps of the CTA keep the
// FP64 pipe busy, so the LENGTH of the dependent chain is what counts.  With  Mh = H / l_w  and  c = b / l_w  (per row, once)
//     u_t = u + beta (u - u_)                        ->  fma(beta, u - u_, u)
//     u_t + (b - H u_g) / l_w                        ->  fma(-Mh_q0, u_g0, fma(-Mh_q1, u_g1, u_t + c))   (u_g = u_t, or u at :163)
// i.e. 5 dependent FP64 operations per iteration instead of 9; the roundings differ from the reference's operation order by
// <= 1 ulp of u per iteration, the same size as the reference's own rounding of u_t + step.
template <bool AT_CURRENT, int NUB>
__device__ __forceinline__ void fused_u_step(const double (&pv)[NUB], const double (&cu)[NUB], const double (&mh)[NUB][NUB],
                                             const double (&c)[NUB], double beta, double (&nx)[NUB]) {
    double ut[NUB];
#pragma unroll
    for (int q = 0; q < NUB; ++q) ut[q] = fma(beta, cu[q] - pv[q], cu[q]);
#pragma unroll
    for (int q = 0; q < NUB; ++q) {
        double un = ut[q] + c[q];
#pragma unroll
        for (int q2 = 0; q2 < NUB; ++q2) un = fma(-mh[q][q2], AT_CURRENT ? cu[q2] : ut[q2], un);
        // np.clip(., 0, 1) by comparisons (a NaN propagates, as in numpy; the simplex projection of the alpha step then stops the fit)
        un = un < 0.0 ? 0.0 : un;
        nx[q] = un > 1.0 ? 1.0 : un;
    }
}
// n2 steps, two at a time so that (u_, u) rotate without register copies; beta_t comes from shared memory
template <bool AT_CURRENT, int NUB>
__device__ __forceinline__ void fused_u_iterate(double (&u)[NUB], double (&up)[NUB], const double (&v)[ng_of(NUB)], uint32_t beta32, int n2,
                                                double inv_lw) {
    double mh[NUB][NUB], c[NUB];
#pragma unroll
    for (int q = 0; q < NUB; ++q) {
        c[q] = v[q] * inv_lw;
#pragma unroll
        for (int q2 = 0; q2 < NUB; ++q2) mh[q][q2] = v[NUB + (q <= q2 ? tri_index(q, q2, NUB) : tri_index(q2, q, NUB))] * inv_lw;
    }
    int itn = 0;
    for (; itn + 2 <= n2; itn += 2) {
        double b0, b1, n1[NUB], n3[NUB];
        lds2(beta32 + (uint32_t)itn * 8u, b0, b1);
        fused_u_step<AT_CURRENT, NUB>(up, u, mh, c, b0, n1);
        fused_u_step<AT_CURRENT, NUB>(u, n1, mh, c, b1, n3);
#pragma unroll
        for (int q = 0; q < NUB; ++q) { up[q] = n1[q]; u[q] = n3[q]; }
    }
    if (itn < n2) {
        double b0, n1[NUB];
        lds1(beta32 + (uint32_t)itn * 8u, b0);
        fused_u_step<AT_CURRENT, NUB>(up, u, mh, c, b0, n1);
#pragma unroll
        for (int q = 0; q < NUB; ++q) { up[q] = u[q]; u[q] = n1[q]; }
    }
}

// A-warps: row statistics of one 8-row block (rb) of a tile for the lane's S samples.  MASK = the tile is short or some of the
// warp's samples do not exist (d = 0 there).  acc[2 rb NG ...] = [b (NUB) | H upper triangle] of row 8 rb + ti, then of row 8 rb + ti + 4.
template <bool MASK, typename WT, int KS, int NUB, int S>
__device__ __forceinline__ void fused_row_block(const FusedArgs& a, uint32_t sb32, int nrows, int rb, int gi, int ti, int K,
                                                const double (&na)[S][KS > 0 ? KS : 1], const double (&au)[S][NUB],
                                                const double (&P)[S][ng_of(NUB) - NUB], const int (&jc)[S], const bool (&valid)[S],
                                                double& cost, double (&acc)[((4 * ng_of(NUB) + 7) / 8) * 8]) {
    constexpr int NG = ng_of(NUB), NTRI = NG - NUB;
    const int o0 = 2 * rb * NG;          // rb is a compile-time constant at both call sites
    const unsigned rpitch = (unsigned)(a.g.ldr * 8);
    const int ra = 8 * rb + ti, rbw = ra + 4;          // the two rows of this lane's C fragment
    const bool la = ra < nrows, lb = rbw < nrows;
    // B operand: R_trunc[row_of_n(gi)][4 kk + ti], MMA column n = 2 t + e  <->  tile row 8 rb + t + 4 e
    double rfrag[KS > 0 ? KS : 1];
    const int rown = 8 * rb + (gi >> 1) + 4 * (gi & 1);
#pragma unroll
    for (int kk = 0; kk < KS; ++kk) {
        const int k = 4 * kk + ti;
        double t = 0.0;
        if (k < K) lds1(sb32 + a.offR + (uint32_t)rown * rpitch + (uint32_t)k * 8u, t);
        rfrag[kk] = t;
    }
    // without masks the lane's samples are jc[0] + 8 sb: constant byte offsets from the first one
    const uint32_t xa = sb32 + (uint32_t)ra * a.pitchX + (MASK ? 0u : (uint32_t)jc[0] * 8u);
    const uint32_t xb = xa + 4u * a.pitchX;
    const uint32_t da = sb32 + a.offD + (uint32_t)ra * a.pitchD + (MASK ? 0u : (uint32_t)jc[0] * (unsigned)sizeof(WT));
    const uint32_t db = da + 4u * a.pitchD;
#pragma unroll
    for (int sb = 0; sb < S; ++sb) {
        const uint32_t ox = MASK ? (uint32_t)jc[sb] * 8u : (uint32_t)sb * 64u;
        const uint32_t od = MASK ? (uint32_t)jc[sb] * (unsigned)sizeof(WT) : (uint32_t)sb * 8u * (unsigned)sizeof(WT);
        double c0, c1;
        lds1(xa + ox, c0);
        lds1(xb + ox, c1);
        double d0 = lds_weight<WT>(da + od);
        double d1 = lds_weight<WT>(db + od);
        if (MASK) {
            if (!valid[sb]) { c0 = 0.0; c1 = 0.0; }
            if (!(valid[sb] && la)) d0 = 0.0;
            if (!(valid[sb] && lb)) d1 = 0.0;
        }
#pragma unroll
        for (int kk = 0; kk < KS; ++kk) dmma884(c0, c1, na[sb][kk], rfrag[kk]);     // c = x - R_trunc a_k
        const double z0 = d0 * c0, z1 = d1 * c1;
        cost = fma(z0, c0, cost);
        cost = fma(z1, c1, cost);
#pragma unroll
        for (int q = 0; q < NUB; ++q) {
            acc[o0 + q] = fma(z0, au[sb][q], acc[o0 + q]);
            acc[o0 + NG + q] = fma(z1, au[sb][q], acc[o0 + NG + q]);
        }
#pragma unroll
        for (int e = 0; e < NTRI; ++e) {
            acc[o0 + NUB + e] = fma(d0, P[sb][e], acc[o0 + NUB + e]);
            acc[o0 + NG + NUB + e] = fma(d1, P[sb][e], acc[o0 + NG + NUB + e]);
        }
    }
}

// C-warps: one tile of the Gram panel with the new u.  B operand column c = 8 nb + gi of  u (x) [R_trunc | u]  is the product
// of two entries of the row's [R_trunc | u] in the stage (offsets fa / fb; columns beyond NCOL point both factors at a zero).
template <bool MASK, typename WT, int NUB, int S, int NBLK>
__device__ __forceinline__ void fused_panel_tile(const FusedArgs& a, uint32_t sb32, int nrows, int ti, const unsigned (&fa_off)[NBLK],
                                                 const unsigned (&fa_pitch)[NBLK], const unsigned (&fb_off)[NBLK], const unsigned (&fb_pitch)[NBLK], const int (&jc)[S],
                                                 const bool (&valid)[S], double (&acc)[S][NBLK][2], double (&accx)[S][NUB]) {
    const unsigned upitch = (unsigned)(a.g.ldu * 8);
#pragma unroll 2
    for (int ks = 0; ks < kFusedRows / 4; ++ks) {
        const int row = 4 * ks + ti;
        const bool lrow = row < nrows;
        double bfrag[NBLK], un[NUB];
#pragma unroll
        for (int nb = 0; nb < NBLK; ++nb) {
            double fa, fb;
            lds1(sb32 + fa_off[nb] + (uint32_t)row * fa_pitch[nb], fa);
            lds1(sb32 + fb_off[nb] + (uint32_t)row * fb_pitch[nb], fb);
            bfrag[nb] = fa * fb;
        }
#pragma unroll
        for (int q = 0; q < NUB; ++q) lds1(sb32 + a.offU + (uint32_t)row * upitch + (uint32_t)q * 8u, un[q]);
        const uint32_t xr = sb32 + (uint32_t)row * a.pitchX + (MASK ? 0u : (uint32_t)jc[0] * 8u);
        const uint32_t dr = sb32 + a.offD + (uint32_t)row * a.pitchD + (MASK ? 0u : (uint32_t)jc[0] * (unsigned)sizeof(WT));
#pragma unroll
        for (int mb = 0; mb < S; ++mb) {
            double x;
            lds1(xr + (MASK ? (uint32_t)jc[mb] * 8u : (uint32_t)mb * 64u), x);
            double d = lds_weight<WT>(dr + (MASK ? (uint32_t)jc[mb] * (unsigned)sizeof(WT) : (uint32_t)mb * 8u * (unsigned)sizeof(WT)));
            if (MASK) {
                if (!(valid[mb] && lrow)) { d = 0.0; x = 0.0; }
            }
            const double dx = d * x;
#pragma unroll
            for (int q = 0; q < NUB; ++q) accx[mb][q] = fma(dx, un[q], accx[mb][q]);
#pragma unroll
            for (int nb = 0; nb < NBLK; ++nb) dmma884(acc[mb][nb][0], acc[mb][nb][1], d, bfrag[nb]);
        }
    }
}

template <typename WT, int KB, int NUB, int S>
__global__ void __launch_bounds__(kFusedThreads, 1) fused_outer_kernel(const FusedArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int KS = (KB + 3) / 4;                  // k-steps of the c = x - R_trunc a_k MMA
    constexpr int NG = ng_of(NUB);
    constexpr int NTRI = NG - NUB;
    constexpr int NCOL = NUB * KB + NTRI;             // panel columns: u_q R_k (q major), then the upper triangle of u u^T
    constexpr int NBLK = (NCOL + 7) / 8;
    constexpr int NS = kFusedStages, TR = kFusedRows;
    constexpr int NV = ((2 * 2 * NG + 7) / 8) * 8;    // per-lane row partials of one tile: 2 row blocks x 2 rows x NG, padded
    const Geom& g = a.g;
    const FitDev f = a.fits[fit_id(g)];
    FitState* st = f.st;
    if (st->done) return;
    FusedCtl* ctl = reinterpret_cast<FusedCtl*>(smem);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int gi = lane >> 2, ti = lane & 3;
    const uint32_t smem32 = smem_u32(smem);
    const uint32_t stages32 = smem32 + kFusedCtlBytes;
    const int part = part_id(g);
    const int n_my = (a.n_tiles > part) ? (a.n_tiles - part + g.n_parts - 1) / g.n_parts : 0;
    const int ucur = st->u_cur, acur = st->a_cur;
    const double* Acur = reinterpret_cast<const double*>(f.A) + (size_t)acur * g.Kt * g.N;
    const int N = g.N, K = g.K;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            mbar_init(smem_u32(&ctl->full[s]), 1);
            mbar_init(smem_u32(&ctl->empty[s]), kFC);
            mbar_init(smem_u32(&ctl->stats[s]), kFA);
            mbar_init(smem_u32(&ctl->udone[s]), 1);
        }
        mbar_init(smem_u32(&ctl->sfree[0]), 1);
        mbar_init(smem_u32(&ctl->sfree[1]), 1);
        mbar_fence_init();
    }
    // every stage starts as finite data: rows beyond the end of the last tile and samples beyond N are weighted with d = 0,
    // which only works on finite values
    {
        const unsigned n16 = (a.offStats - kFusedCtlBytes) / 16;
        for (unsigned i = tid; i < n16; i += kFusedThreads)
            asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(stages32 + i * 16u), "r"(0u) : "memory");
    }
    if (tid < a.n_iter2) {
        // beta_t = min((a_t - 1) / a_{t+1}, 0.9999 sqrt(l_w_ / l_w)); l_w_ == l_w from the second inner iteration on (:89)
        const double l_w = st->l_w;
        const double cap = 0.9999 * sqrt((tid == 0 ? st->l_w_old : l_w) / l_w);
        ctl->beta[tid] = fmin(a.mom_m[st->t_u + tid], cap);
    }
    fence_proxy_async_smem();
    __syncthreads();

    auto tile_rows = [&](int it) {
        const long long r0 = ((long long)part + (long long)it * g.n_parts) * TR;
        const long long left = g.M - r0;
        return (int)(left < TR ? left : TR);
    };

    double cost = 0.0, ssq = 0.0;       // per-thread partials (A-warps: sum d c^2; U-warps: cross terms, ||u_new||^2)
    // C-warp state (declared here because it is stored after the CTA-wide barrier that follows the role loops)
    double pacc[S][NBLK][2], paccx[S][NUB];
    int jcC[S];
    bool validC[S];

    if (warp >= kFA + kFC + kFU) {
        // =========================================================================== producer warp: the TMA ring
        asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
        const char* Ucur_g = f.U + (size_t)ucur * g.uslot_bytes;
        const char* Uprv_g = f.U + (size_t)(ucur ^ 1) * g.uslot_bytes;
        const unsigned rbytes = (unsigned)(g.ldr * 8), ubytes = (unsigned)(g.ldu * 8);
        for (int it = 0; it < n_my; ++it) {
            const int s = it % NS;
            mbar_wait(smem_u32(&ctl->empty[s]), (((unsigned)(it / NS)) & 1u) ^ 1u);      // the C-warps released tile it - NS
            const long long r0 = ((long long)part + (long long)it * g.n_parts) * TR;
            const int nrows = tile_rows(it);
            const uint32_t full = smem_u32(&ctl->full[s]);
            const uint32_t sb = stages32 + (uint32_t)s * a.stage_bytes;
            if (lane == 0) {
                fence_proxy_async_smem();      // the U-warps wrote the new u into this stage through the generic proxy
                mbar_arrive_expect_tx(full, (unsigned)nrows * (a.pitchX + a.pitchD + rbytes + 2u * ubytes));
            }
            __syncwarp();
            // one bulk copy per matrix: a tile is contiguous in global memory and keeps its row pitch in shared memory (the caller
            // pads the rows of X and d_x so that the pitch is bank-conflict free, see dmf_shape_t)
            if (lane == 0) bulk_g2s(sb, f.X + r0 * (long long)a.pitchX, (unsigned)nrows * a.pitchX, full);
            if (lane == 1) bulk_g2s(sb + a.offD, f.D + r0 * (long long)a.pitchD, (unsigned)nrows * a.pitchD, full);
            if (lane == 2 && K) bulk_g2s(sb + a.offR, f.Rk + r0 * (long long)rbytes, (unsigned)nrows * rbytes, full);
            if (lane == 3) bulk_g2s(sb + a.offU, Ucur_g + r0 * (long long)ubytes, (unsigned)nrows * ubytes, full);
            if (lane == 4) bulk_g2s(sb + a.offUp, Uprv_g + r0 * (long long)ubytes, (unsigned)nrows * ubytes, full);
            __syncwarp();
        }
    } else if (warp >= kFA + kFC) {
        // =========================================================================== U-warps: update_u on the row statistics
        asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
        const int uw = warp - (kFA + kFC);
        double* Unew_g = reinterpret_cast<double*>(f.U + (size_t)(ucur ^ 2) * g.uslot_bytes);
        double* Unpv_g = reinterpret_cast<double*>(f.U + (size_t)(ucur ^ 3) * g.uslot_bytes);
        const int n2 = a.n_iter2;
        const double l_w = st->l_w;
        const double inv_lw = 1.0 / l_w;       // as u_inner_kernel: reciprocal multiply (<= 1 ulp of a ~1e-4-sized step vs the division of :88)
        const bool at_current = (g.mode == 2);
        const unsigned upitch = (unsigned)(g.ldu * 8);
        // beta_t = min((a_t - 1) / a_{t+1}, 0.9999 sqrt(l_w_ / l_w)); l_w_ == l_w from the second inner iteration on (:89)
        const uint32_t beta32 = smem_u32(&ctl->beta[0]);
        for (int it = uw; it < n_my; it += kFU) {
            const int s = it % NS;
            const unsigned ph = ((unsigned)(it / NS)) & 1u;
            mbar_wait(smem_u32(&ctl->stats[s]), ph);
            mbar_wait(smem_u32(&ctl->full[s]), ph);        // complete long ago; orders this warp after the bulk copies
            const int nrows = tile_rows(it);
            const long long r0 = ((long long)part + (long long)it * g.n_parts) * TR;
            const uint32_t sb = stages32 + (uint32_t)s * a.stage_bytes;
            const bool live = lane < nrows;
            const int row = lane < TR ? lane : 0;
            // row statistics: the 8 A-warp partials in warp order; the buffer is free for tile it + 2 as soon as they are in registers
            double v[NG];
#pragma unroll
            for (int i = 0; i < NG; ++i) v[i] = 0.0;
            {
                const uint32_t sbase = smem32 + a.offStats + (uint32_t)(it & 1) * (kFA * TR * NG * 8u) + (uint32_t)row * (NG * 8u);
#pragma unroll
                for (int w = 0; w < kFA; ++w)
#pragma unroll
                    for (int i = 0; i < NG; ++i) {
                        double t;
                        lds1(sbase + (uint32_t)w * (TR * NG * 8u) + (uint32_t)i * 8u, t);
                        v[i] += t;
                    }
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&ctl->sfree[it & 1]));
            }
            double u[NUB], up[NUB];
#pragma unroll
            for (int q = 0; q < NUB; ++q) {
                lds1(sb + a.offU + (uint32_t)row * upitch + (uint32_t)q * 8u, u[q]);
                lds1(sb + a.offUp + (uint32_t)row * upitch + (uint32_t)q * 8u, up[q]);
            }
            if (live) {      // cost of the incoming iterate: -2 u^T b + u^T H u of this row (rowgram4_kernel, writer lanes)
                double ct = 0.0;
#pragma unroll
                for (int q = 0; q < NUB; ++q) {
                    double hq = 0.0;
#pragma unroll
                    for (int q2 = 0; q2 < NUB; ++q2) hq = fma(v[NUB + (q <= q2 ? tri_index(q, q2, NUB) : tri_index(q2, q, NUB))], u[q2], hq);
                    ct = fma(u[q], hq - 2.0 * v[q], ct);
                }
                cost += ct;
            }
            if (at_current) fused_u_iterate<true, NUB>(u, up, v, beta32, n2, inv_lw);
            else fused_u_iterate<false, NUB>(u, up, v, beta32, n2, inv_lw);
            // new iterate: global (other slot pair) and the stage (the C-warps form u (x) [R_trunc | u] and bx_u from it)
            if (live) {
#pragma unroll
                for (int q = 0; q < NUB; ++q) {
                    if (q < g.nu) {
                        Unew_g[(size_t)(r0 + lane) * g.ldu + q] = u[q];
                        Unpv_g[(size_t)(r0 + lane) * g.ldu + q] = up[q];
                        ssq = fma(u[q], u[q], ssq);
                    }
                }
            }
            if (lane < TR) {
#pragma unroll
                for (int q = 0; q < NUB; ++q) {
                    const double un = (live && q < g.nu) ? u[q] : 0.0;
                    asm volatile("st.shared.f64 [%0], %1;" ::"r"(sb + a.offU + (uint32_t)lane * upitch + (uint32_t)q * 8u), "d"(un) : "memory");
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&ctl->udone[s]));
        }
    } else if (warp < kFA) {
        // =========================================================================== A-warps: row statistics
        asm volatile("setmaxnreg.inc.sync.aligned.u32 112;");
        double na[S][KS > 0 ? KS : 1], au[S][NUB], P[S][NTRI > 0 ? NTRI : 1];
        int jc[S];
        bool valid[S];
#pragma unroll
        for (int sb = 0; sb < S; ++sb) {
            const int j = 8 * (warp * S + sb) + gi;
            valid[sb] = j < N;
            jc[sb] = valid[sb] ? j : 0;
#pragma unroll
            for (int kk = 0; kk < KS; ++kk) {
                const int k = 4 * kk + ti;
                na[sb][kk] = (valid[sb] && k < K) ? -Acur[(size_t)k * N + j] : 0.0;
            }
#pragma unroll
            for (int q = 0; q < NUB; ++q) au[sb][q] = (valid[sb] && q < g.nu) ? Acur[(size_t)(K + q) * N + j] : 0.0;
#pragma unroll
            for (int q = 0; q < NUB; ++q)
#pragma unroll
                for (int q2 = q; q2 < NUB; ++q2) P[sb][tri_index(q, q2, NUB)] = au[sb][q] * au[sb][q2];
        }
        const unsigned rpitch = (unsigned)(g.ldr * 8);
        const bool cols_full = 8 * (warp * S + S) <= N;       // every sample of this warp exists: no column masks
        for (int it = 0; it < n_my; ++it) {
            const int s = it % NS;
            const unsigned ph = ((unsigned)(it / NS)) & 1u;
            mbar_wait(smem_u32(&ctl->full[s]), ph);
            if (it >= 2) mbar_wait(smem_u32(&ctl->sfree[it & 1]), ((unsigned)((it >> 1) - 1)) & 1u);     // tile it - 2's partials were read
            const int nrows = tile_rows(it);
            const uint32_t sb32 = stages32 + (uint32_t)s * a.stage_bytes;
            double acc[NV];
#pragma unroll
            for (int i = 0; i < NV; ++i) acc[i] = 0.0;
            if (cols_full && nrows == TR) {
                fused_row_block<false, WT, KS, NUB, S>(a, sb32, nrows, 0, gi, ti, K, na, au, P, jc, valid, cost, acc);
                fused_row_block<false, WT, KS, NUB, S>(a, sb32, nrows, 1, gi, ti, K, na, au, P, jc, valid, cost, acc);
            } else {
                fused_row_block<true, WT, KS, NUB, S>(a, sb32, nrows, 0, gi, ti, K, na, au, P, jc, valid, cost, acc);
                fused_row_block<true, WT, KS, NUB, S>(a, sb32, nrows, 1, gi, ti, K, na, au, P, jc, valid, cost, acc);
            }
            // sum over the 8 samples of a block row (lanes sharing ti), then one partial per (warp, row, value)
            const int base = reduce_over_groups<NV>(acc, lane);
            const uint32_t sbase = smem32 + a.offStats + (uint32_t)(it & 1) * (kFA * TR * NG * 8u) + (uint32_t)warp * (TR * NG * 8u);
#pragma unroll
            for (int i = 0; i < NV / 8; ++i) {
                const int slot = base + i;
                if (slot < 4 * NG) {
                    const int re = slot / NG, vi = slot - re * NG;       // re = rb * 2 + e
                    const int row = 8 * (re >> 1) + ti + 4 * (re & 1);
                    asm volatile("st.shared.f64 [%0], %1;" ::"r"(sbase + (uint32_t)(row * NG + vi) * 8u), "d"(acc[i]) : "memory");
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&ctl->stats[s]));
        }
    } else {
        // =========================================================================== C-warps: Gram panel with the new u
        const int cw = warp - kFA;
#pragma unroll
        for (int mb = 0; mb < S; ++mb) {
            const int j = 8 * (cw * S + mb) + gi;
            validC[mb] = j < N;
            jcC[mb] = validC[mb] ? j : 0;
#pragma unroll
            for (int nb = 0; nb < NBLK; ++nb) { pacc[mb][nb][0] = 0.0; pacc[mb][nb][1] = 0.0; }
#pragma unroll
            for (int q = 0; q < NUB; ++q) paccx[mb][q] = 0.0;
        }
        const unsigned upitch = (unsigned)(g.ldu * 8), rpitch = (unsigned)(g.ldr * 8);
        // B operand of the panel MMA: column c = 8 nb + gi of  u (x) [R_trunc | u]  =  (first factor u_q) x (second factor R_k or u_q2),
        // formed per k-step from the stage (the U-warp left the new u there)
        unsigned fa_off[NBLK], fa_pitch[NBLK], fb_off[NBLK], fb_pitch[NBLK];
        const unsigned zero_off = a.zero_off;        // 8 bytes of every stage that are zeroed at kernel start and never copied over
#pragma unroll
        for (int nb = 0; nb < NBLK; ++nb) {
            const int c = 8 * nb + gi;
            int q = 0, q2 = 0, k = -1;
            if (c < NUB * KB) { q = c / (KB > 0 ? KB : 1); k = c - q * KB; }
            else if (c < NCOL) {
                int e2 = c - NUB * KB;
                while (e2 >= NUB - q) { e2 -= NUB - q; ++q; }
                q2 = q + e2;
            }
            const bool fzero = c >= NCOL || q >= g.nu || (k >= 0 ? k >= K : q2 >= g.nu);
            // a column that does not exist multiplies a zero by itself (no select in the loop)
            fa_off[nb] = fzero ? zero_off : a.offU + (unsigned)q * 8u;
            fa_pitch[nb] = fzero ? 0u : upitch;
            fb_off[nb] = fzero ? zero_off : ((k >= 0) ? a.offR + (unsigned)k * 8u : a.offU + (unsigned)q2 * 8u);
            fb_pitch[nb] = fzero ? 0u : ((k >= 0) ? rpitch : upitch);
        }
        const bool cols_full = 8 * (cw * S + S) <= N;
        for (int it = 0; it < n_my; ++it) {
            const int s = it % NS;
            const unsigned ph = ((unsigned)(it / NS)) & 1u;
            mbar_wait(smem_u32(&ctl->udone[s]), ph);
            mbar_wait(smem_u32(&ctl->full[s]), ph);
            const int nrows = tile_rows(it);
            const uint32_t sb32 = stages32 + (uint32_t)s * a.stage_bytes;
            if (cols_full && nrows == TR) fused_panel_tile<false, WT, NUB, S, NBLK>(a, sb32, nrows, ti, fa_off, fa_pitch, fb_off, fb_pitch, jcC, validC, pacc, paccx);
            else fused_panel_tile<true, WT, NUB, S, NBLK>(a, sb32, nrows, ti, fa_off, fa_pitch, fb_off, fb_pitch, jcC, validC, pacc, paccx);
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&ctl->empty[s]));
        }
    }
    __syncthreads();
    if (warp >= kFA && warp < kFA + kFC) {
        const int cw = warp - kFA;
        // CTA record: [cost, ||u||^2, panel NCOL x N, bx NUB x N] - the stage ring is free once every role left its loop
        double* rec = reinterpret_cast<double*>(smem + kFusedCtlBytes);
#pragma unroll
        for (int mb = 0; mb < S; ++mb) {
            const int j = 8 * (cw * S + mb) + gi;
#pragma unroll
            for (int nb = 0; nb < NBLK; ++nb)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int c = 8 * nb + 2 * ti + e;
                    if (validC[mb] && c < NCOL) rec[2 + (size_t)c * N + j] = pacc[mb][nb][e];
                }
#pragma unroll
            for (int q = 0; q < NUB; ++q) {
                double t = paccx[mb][q];
                t += __shfl_xor_sync(0xffffffffu, t, 1);
                t += __shfl_xor_sync(0xffffffffu, t, 2);
                if (validC[mb] && ti == 0) rec[2 + (size_t)(NCOL + q) * N + j] = t;
            }
        }
    }
    // scalar partials: warp sums, then the warps in order
    {
        double c = cost, q = ssq;
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) {
            c += __shfl_xor_sync(0xffffffffu, c, o);
            q += __shfl_xor_sync(0xffffffffu, q, o);
        }
        if (lane == 0) { ctl->wsum[0][warp] = c; ctl->wsum[1][warp] = q; }
    }
    __syncthreads();
    double* rec = reinterpret_cast<double*>(smem + kFusedCtlBytes);
    if (tid == 0) {
        double c = 0.0, q = 0.0;
        for (int w = 0; w < kFA + kFC + kFU + kFP; ++w) { c += ctl->wsum[0][w]; q += ctl->wsum[1][w]; }
        rec[0] = c;
        rec[1] = q;
    }
    __syncthreads();
    const int n = 2 + (NCOL + NUB) * N;
    {
        double* p = f.part + (size_t)part * g.part_stride;
        for (int e = tid; e < n; e += kFusedThreads) p[e] = rec[e];
    }
    if (!hier_reduce(g, f, f.red, n, &ctl->flag)) return;
    __threadfence();
    __syncthreads();
    // ------------------------------------------------------------------------------- last CTA of the fit: test, then commit
    if (tid == 0) {
        const double cf = f.red[0], su = f.red[1];
        int commit = 1;
        if (a.flags & kFlagPartial) {
            // CpG rows sharded over GPUs: publish this GPU's sums; alpha_inner_kernel tests and commits on the all-reduced sums
            f.scal[0] = cf;
            f.scal[4] = su;
        } else {
            if (st->phase == 1) {              // the incoming iterate closes an outer iteration: deconvolution.py:218-221
                const double prev = st->cf;
                st->cf_prev = prev;
                st->cf = cf;
                const int no = st->n_outer + 1;
                st->n_outer = no;
                if (f.trace && no < f.trace_cap) f.trace[no] = cf;
                st->phase = 0;
                if (fabs(cf - prev) < a.tol) { st->done = 1; commit = 0; }
                if (!(cf == cf)) { st->done = 3; commit = 0; }
            }
            if (commit) {
                const int t0 = st->t_u, n2 = a.n_iter2;
                st->u_cur = ucur ^ 2;
                st->a1 = a.mom_a[t0 + n2];
                st->t_u = t0 + n2;
                if (n2 > 0) st->l_w_old = st->l_w;               // deconvolution.py:89
                st->ssq_u = su;
                const double nr = sqrt(st->ssq_rk + su);
                st->l_h = (nr * nr) * st->dmax2;                 // deconvolution.py:212
                st->phase = 1;
            }
        }
        ctl->commit = commit;
    }
    __syncthreads();
    if (!ctl->commit) return;
    for (int e = tid; e < (NCOL + NUB) * N; e += kFusedThreads) {
        const int c = e / N, j = e - c * N;
        const double v = f.red[2 + e];
        if (c >= NCOL) {
            const int q = c - NCOL;
            if (q < g.nu) f.gbx[(size_t)(K + q) * N + j] = v;
        } else if (c < NUB * KB) {
            const int q = c / KB, k = c - q * KB;
            if (q < g.nu && k < K) {
                f.gram[((size_t)(K + q) * g.Kt + k) * N + j] = v;
                f.gram[((size_t)k * g.Kt + (K + q)) * N + j] = v;
            }
        } else {
            int e2 = c - NUB * KB, q = 0;
            while (e2 >= NUB - q) { e2 -= NUB - q; ++q; }
            const int q2 = q + e2;
            if (q2 < g.nu) {
                f.gram[((size_t)(K + q) * g.Kt + (K + q2)) * N + j] = v;
                f.gram[((size_t)(K + q2) * g.Kt + (K + q)) * N + j] = v;
            }
        }
    }
}

}  // namespace dmf
