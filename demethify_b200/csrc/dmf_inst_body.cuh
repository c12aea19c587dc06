// dmf_inst_body.cuh — included by dmf_inst_<tag>.cu with DMF_T, DMF_WT and DMF_TAG defined.
#include "dmf_inst.h"
#include "dmf_kernels.cuh"
#include "dmf_wls.cuh"
#include "dmf_gram.cuh"
#include "dmf_fused.cuh"
namespace dmf {
#define DMF_CAT2(a, b) a##b
#define DMF_CAT(a, b) DMF_CAT2(a, b)
// cost / set-up pass: (register-row bucket, columns per thread)
kern_t DMF_CAT(pick_cost_, DMF_TAG)(int ktb, int initial, int c) {
#define DMF_C(KTB_, C_)                                                                              \
    if (ktb == KTB_ && c == C_)                                                                      \
        return initial ? (kern_t)cost_kernel<DMF_T, DMF_WT, KTB_, C_, true> : (kern_t)cost_kernel<DMF_T, DMF_WT, KTB_, C_, false>;
    DMF_C(8, 2) DMF_C(16, 2) DMF_C(32, 1)
#undef DMF_C
    return nullptr;
}
kern_t DMF_CAT(pick_alpha_, DMF_TAG)(int ktb, int, int c) {
    if (ktb == 8 && c == 2) return alpha_pass_kernel<DMF_T, DMF_WT, 8, 2>;
    if (ktb == 16 && c == 2) return alpha_pass_kernel<DMF_T, DMF_WT, 16, 2>;
    if (ktb == 32 && c == 1) return alpha_pass_kernel<DMF_T, DMF_WT, 32, 1>;
    return nullptr;
}
// U pass: (known bucket, unknown bucket) -> fixed (columns per thread, rows per thread per tile); see kUTable in dmf_api.cu
kern_t DMF_CAT(pick_u_, DMF_TAG)(int kb, int nub, int) {
#define DMF_U(KB_, NUB_, C_, RPT_) \
    if (kb == KB_ && nub == NUB_) return u_pass_kernel<DMF_T, DMF_WT, KB_, NUB_, C_, RPT_>;
    DMF_U(6, 2, 2, 4) DMF_U(8, 2, 2, 4) DMF_U(8, 8, 2, 1) DMF_U(16, 2, 2, 4) DMF_U(16, 8, 2, 1) DMF_U(16, 16, 1, 1)
    DMF_U(32, 2, 1, 4) DMF_U(32, 8, 1, 1) DMF_U(32, 32, 1, 1)
#undef DMF_U
    return nullptr;
}
// Gram engine (dmf_gram.cuh).  rowgram: (known bucket, unknown bucket) -> fixed (columns per thread, rows per thread per tile);
// see kGramTable in dmf_api.cu
kern_t DMF_CAT(pick_rowgram_, DMF_TAG)(int kb, int nub, int flags) {
    const int initial = flags & 1, c4 = flags & 2;
#define DMF_G4(KB_, NUB_, RPT_)                                                                              \
    if (c4 && kb == KB_ && nub == NUB_)                                                                      \
        return initial ? (kern_t)rowgram4_kernel<DMF_T, DMF_WT, KB_, NUB_, RPT_, true> : (kern_t)rowgram4_kernel<DMF_T, DMF_WT, KB_, NUB_, RPT_, false>;
    DMF_G4(0, 1, 4) DMF_G4(0, 2, 4) DMF_G4(0, 4, 2) DMF_G4(0, 8, 1) DMF_G4(6, 1, 4) DMF_G4(6, 2, 4) DMF_G4(6, 4, 2) DMF_G4(6, 8, 1)
#undef DMF_G4
    if (c4) return nullptr;
#define DMF_G(KB_, NUB_, C_, RPT_)                                                                           \
    if (kb == KB_ && nub == NUB_)                                                                            \
        return initial ? (kern_t)rowgram_kernel<DMF_T, DMF_WT, KB_, NUB_, C_, RPT_, true> : (kern_t)rowgram_kernel<DMF_T, DMF_WT, KB_, NUB_, C_, RPT_, false>;
    DMF_G(0, 1, 2, 4) DMF_G(0, 2, 2, 3) DMF_G(0, 4, 2, 2)
    DMF_G(6, 1, 2, 4) DMF_G(6, 2, 2, 3) DMF_G(6, 4, 2, 2)
    DMF_G(16, 1, 2, 4) DMF_G(16, 2, 2, 3) DMF_G(16, 4, 2, 2)
    DMF_G(32, 1, 1, 4) DMF_G(32, 2, 1, 3) DMF_G(32, 4, 1, 2)
#undef DMF_G
    return nullptr;
}
// mult: 0 plain, 1 multiplicity form, 2 multiplicity form, u block of a single unknown type (PA = 1)
kern_t DMF_CAT(pick_panel_, DMF_TAG)(int pb, int c, int mult) {
    if (mult == 2) return (pb == 8 && c == 4) ? (kern_t)gram_panel_kernel<DMF_T, DMF_WT, 1, 8, 4, true> : nullptr;
    if (mult) return (pb == 8 && c == 4) ? (kern_t)gram_panel_kernel<DMF_T, DMF_WT, 2, 8, 4, true> : nullptr;
    if (pb == 8 && c == 4) return gram_panel_kernel<DMF_T, DMF_WT, 2, 8, 4, false>;
    if (pb == 8) return gram_panel_kernel<DMF_T, DMF_WT, 2, 8, 2, false>;
    if (pb == 16) return gram_panel_kernel<DMF_T, DMF_WT, 2, 16, 1, false>;
    return nullptr;
}
// which: 0 u_inner_kernel, 1 u_inner_mult_kernel, 2 cost_cross_kernel, 3 usum_kernel
kern_t DMF_CAT(pick_uinner_, DMF_TAG)(int nub, int which, int) {
#define DMF_UI(NUB_)                                                           \
    if (nub == NUB_) {                                                         \
        if (which == 0) return u_inner_kernel<DMF_T, NUB_>;                    \
        if (which == 1) return u_inner_mult_kernel<DMF_T, NUB_>;               \
        if (which == 3) return usum_kernel<DMF_T, NUB_>;                       \
        return cost_cross_kernel<DMF_T, NUB_>;                                 \
    }
    DMF_UI(1) DMF_UI(2) DMF_UI(4) DMF_UI(8)
#undef DMF_UI
    return nullptr;
}
kern_t DMF_CAT(pick_ainner_, DMF_TAG)(int ktb, int, int) {
    if (ktb == 8) return alpha_inner_kernel<DMF_T, 8>;
    if (ktb == 16) return alpha_inner_kernel<DMF_T, 16>;
    if (ktb == 32) return alpha_inner_kernel<DMF_T, 32>;
    return nullptr;
}
// fused engine (dmf_fused.cuh): FP64 storage only; (known bucket, unknown types, 8-sample blocks per warp)
template <typename T, typename WT>
struct FusedPick {
    static fused_kern_t get(int, int, int) { return nullptr; }
};
template <typename WT>
struct FusedPick<double, WT> {
    static fused_kern_t get(int kb, int nub, int s) {
#define DMF_F(KB_, NUB_, S_) \
    if (kb == KB_ && nub == NUB_ && s == S_) return fused_outer_kernel<WT, KB_, NUB_, S_>;
#define DMF_FS(KB_, NUB_) DMF_F(KB_, NUB_, 1) DMF_F(KB_, NUB_, 2) DMF_F(KB_, NUB_, 4)
        DMF_FS(0, 1) DMF_FS(0, 2) DMF_FS(4, 1) DMF_FS(4, 2) DMF_FS(6, 1) DMF_FS(6, 2) DMF_FS(8, 1) DMF_FS(8, 2)
#undef DMF_FS
#undef DMF_F
        return nullptr;
    }
};
fused_kern_t DMF_CAT(pick_fused_, DMF_TAG)(int kb, int nub, int s) { return FusedPick<DMF_T, DMF_WT>::get(kb, nub, s); }
wls_kern_t DMF_CAT(pick_wls_, DMF_TAG)() { return wls_moments_kernel<DMF_T, DMF_WT>; }
}  // namespace dmf
