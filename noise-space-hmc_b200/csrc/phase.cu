// Phase-retrieval operator (Hfuncs.py:318-366): placeholder until the shared-memory FFT lands.
#include "common.cuh"
namespace nhmc {
size_t phase_workspace_bytes(const nhmc_op*, int) { return 256; }
int phase_forward(const nhmc_op*, const float*, float*, int, void*, cudaStream_t) { set_error("phase: not built"); return NHMC_ERR_UNSUPPORTED; }
int phase_pinv(const nhmc_op*, const float*, float*, int, void*, cudaStream_t) { set_error("phase: not built"); return NHMC_ERR_UNSUPPORTED; }
int phase_loss_grad(const nhmc_op*, const float*, const float*, float*, float*, double*, int, void*, cudaStream_t) { set_error("phase: not built"); return NHMC_ERR_UNSUPPORTED; }
}
