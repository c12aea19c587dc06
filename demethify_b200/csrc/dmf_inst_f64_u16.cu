// kernel instantiations for X/R/U/alpha = double, d_x stored as uint16_t
#include <cstdint>
#define DMF_T double
#define DMF_WT uint16_t
#define DMF_TAG f64_u16
#include "dmf_inst_body.cuh"
