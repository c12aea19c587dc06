// dmf_inst.h — kernel lookup functions, one set per (element type, weight storage) translation unit.
#pragma once
#include "dmf_device.cuh"
namespace dmf {
typedef void (*kern_t)(const PassArgs);
struct FusedArgs;
typedef void (*fused_kern_t)(const FusedArgs);
struct WlsArgs;
typedef void (*wls_kern_t)(const WlsArgs);
#define DMF_DECL(TAG)                                      \
    kern_t pick_cost_##TAG(int ktb, int nub, int c);       \
    kern_t pick_alpha_##TAG(int ktb, int nub, int c);      \
    kern_t pick_u_##TAG(int ktb, int nub, int c);          \
    kern_t pick_rowgram_##TAG(int kb, int nub, int initial); \
    kern_t pick_panel_##TAG(int pb, int, int);             \
    kern_t pick_uinner_##TAG(int nub, int, int);           \
    kern_t pick_ainner_##TAG(int ktb, int, int);           \
    fused_kern_t pick_fused_##TAG(int kb, int nub, int s);   \
    wls_kern_t pick_wls_##TAG();
DMF_DECL(f64_f64) DMF_DECL(f64_u16) DMF_DECL(f32_f32) DMF_DECL(f32_u16)
#undef DMF_DECL
}  // namespace dmf
