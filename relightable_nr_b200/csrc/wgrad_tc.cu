// tcgen05 weight-gradient kernel (MN-major operands).  See DESIGN.md "wgrad".
#include "conv_internal.cuh"

int rnr_wgrad_tc_prepare(rnr_wgrad_plan* plan, const rnr_wgrad_problem_t* prob) {
    (void)plan; (void)prob;
    rnr_set_error("rnr_wgrad: tcgen05 implementation not available in this build");
    return (int)cudaErrorNotSupported;
}
int rnr_wgrad_tc_run(const rnr_wgrad_plan* plan, cudaStream_t stream) {
    (void)plan; (void)stream;
    rnr_set_error("rnr_wgrad: tcgen05 implementation not available in this build");
    return (int)cudaErrorNotSupported;
}
