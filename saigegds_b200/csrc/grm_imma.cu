// int8 tensor-core (Ozaki-sliced, exact integer) implementation of get_crossprod_b_grm.
// Placeholder until the kernel lands: reports "not available" so that SGB_KERNEL_AUTO uses the
// FP64 CUDA-core path; asking for SGB_KERNEL_IMMA explicitly is an error, never a silent fallback.
#include "ctx.h"

namespace sgb {

void imma_prepare(Context &) {}
void imma_release(Context &) {}
bool imma_available(const Context &) { return false; }
void imma_grm_mv(Context &, const double *, double *, int) {
    throw Error(SGB_ERR_STATE, "the IMMA product kernel is not built in this version");
}

}  // namespace sgb
