// tcgen05 / TMEM engine of the fused sparse step — placeholder until the 3xTF32 kernel lands.
#include "common.cuh"

namespace lrs {

bool sparse_fused_tc_supported(const FusedParams&, int) { return false; }

int sparse_fused_tc_launch(const FusedParams&, int, cudaStream_t) {
    return fail_arg("lrs_sparse_step_fused_f32", "tcgen05 engine not built");
}

}  // namespace lrs
