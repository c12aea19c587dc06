// kernel instantiations for X/R/U/alpha = float, d_x stored as uint16_t
#include <cstdint>
#define DMF_T float
#define DMF_WT uint16_t
#define DMF_TAG f32_u16
#include "dmf_inst_body.cuh"
