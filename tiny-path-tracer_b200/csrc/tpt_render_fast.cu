// csrc/tpt_render_fast.cu -- FAST instantiation. Compiled with -use_fast_math (FMA contraction,
// approximate div/sqrt/sin/cos, flush-to-zero): same estimator, same Philox stream, fp32 only.
#define TPT_PAR false
#define TPT_SUFFIX fast
#include "tpt_kernels.cuh"
