// dmf_inst_body.cuh — included by dmf_inst_<tag>.cu with DMF_T, DMF_WT and DMF_TAG defined.
#include "dmf_inst.h"
#include "dmf_kernels.cuh"
namespace dmf {
#define DMF_CAT2(a, b) a##b
#define DMF_CAT(a, b) DMF_CAT2(a, b)
kern_t DMF_CAT(pick_cost_, DMF_TAG)(int ktb, int, int c) {
    if (ktb == 8 && c == 2) return init_cost_kernel<DMF_T, DMF_WT, 8, 2>;
    if (ktb == 8 && c == 4) return init_cost_kernel<DMF_T, DMF_WT, 8, 4>;
    if (ktb == 16 && c == 1) return init_cost_kernel<DMF_T, DMF_WT, 16, 1>;
    if (ktb == 16 && c == 2) return init_cost_kernel<DMF_T, DMF_WT, 16, 2>;
    if (ktb == 32 && c == 1) return init_cost_kernel<DMF_T, DMF_WT, 32, 1>;
    return nullptr;
}
kern_t DMF_CAT(pick_alpha_, DMF_TAG)(int ktb, int, int c) {
    if (ktb == 8 && c == 2) return alpha_pass_kernel<DMF_T, DMF_WT, 8, 2>;
    if (ktb == 8 && c == 4) return alpha_pass_kernel<DMF_T, DMF_WT, 8, 4>;
    if (ktb == 16 && c == 1) return alpha_pass_kernel<DMF_T, DMF_WT, 16, 1>;
    if (ktb == 16 && c == 2) return alpha_pass_kernel<DMF_T, DMF_WT, 16, 2>;
    if (ktb == 32 && c == 1) return alpha_pass_kernel<DMF_T, DMF_WT, 32, 1>;
    return nullptr;
}
kern_t DMF_CAT(pick_u_, DMF_TAG)(int ktb, int nub, int c) {
#define DMF_U(KTB_, NUB_, C_) \
    if (ktb == KTB_ && nub == NUB_ && c == C_) return u_pass_kernel<DMF_T, DMF_WT, KTB_, NUB_, C_>;
    DMF_U(8, 2, 2) DMF_U(8, 2, 4) DMF_U(8, 8, 2) DMF_U(8, 8, 4)
    DMF_U(16, 2, 1) DMF_U(16, 2, 2) DMF_U(16, 8, 1) DMF_U(16, 8, 2) DMF_U(16, 16, 1) DMF_U(16, 16, 2)
    DMF_U(32, 2, 1) DMF_U(32, 8, 1) DMF_U(32, 32, 1)
#undef DMF_U
    return nullptr;
}
}  // namespace dmf
