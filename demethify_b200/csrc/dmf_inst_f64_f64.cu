// kernel instantiations for X/R/U/alpha = double, d_x stored as double
#include <cstdint>
#define DMF_T double
#define DMF_WT double
#define DMF_TAG f64_f64
#include "dmf_inst_body.cuh"
