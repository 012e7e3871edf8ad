// Tuned FP32 step kernel (placeholder: forwards to the basic FP32 kernel until the tuned variant lands).
#include "wf_device.cuh"

cudaError_t wf_launch_step_fast(int mode, const WfModel& m, const WfState& s, const uint8_t* d_mask,
                                const float* d_action, const double* d_yaw_cmd, const WfOutPtrs& out, int sm_count,
                                cudaStream_t stream) {
    (void)sm_count;
    return wf_launch_step_basic(1, mode, m, s, d_mask, d_action, d_yaw_cmd, out, stream);
}

cudaError_t wf_step_fast_attributes(const WfModel& m, cudaFuncAttributes* attr, int* ctas_per_sm, int* threads,
                                    int* smem) {
    *threads = (m.T + 31) / 32 * 32;
    cudaError_t e = wf_step_basic_attributes(1, attr, ctas_per_sm, *threads);
    *smem = (int)attr->sharedSizeBytes;
    return e;
}
