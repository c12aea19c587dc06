// kernel instantiations for X/R/U/alpha = float, d_x stored as float
#include <cstdint>
#define DMF_T float
#define DMF_WT float
#define DMF_TAG f32_f32
#include "dmf_inst_body.cuh"
