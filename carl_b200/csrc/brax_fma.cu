// The FMA-contracted build of the Brax kernels (see the head of brax.cu): same source, compiled WITHOUT
// -fmad=false, in its own inner namespace. Selected per handle with carlb_brax_set_arithmetic.
#define CARLB_BRAX_FMA_BUILD 1
#include "brax.cu"
