// cloud_march_fma.cu -- K1 / K1p / K1s under the CONTRACTED arithmetic definition (MM_ARITH_FMA): cloud_march.cu compiled a second
// time with every multiply-add of the shader fused (see the header of cloud_march.cu, "ARITHMETIC DEFINITIONS").  Exports
// launch_cloud_march_fma and launch_det_pow_fma; the passes that exist once (cloud shadows, probes, self-tests) stay in cloud_march.cu.
#define MM_FMA 1
#include "cloud_march.cu"
