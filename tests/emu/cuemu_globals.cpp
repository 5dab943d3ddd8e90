// TEST-ONLY: storage for the host emulation of the CUDA launch model (see gnark_symmetric_crypto_b200/csrc/cuemu.h).
#include "cuemu.h"
#include <chrono>
namespace cuemu {
thread_local dim3 t_threadIdx, t_blockIdx;
dim3 g_blockDim, g_gridDim;
std::barrier<>* g_barrier = nullptr;
unsigned char* g_dyn_smem = nullptr;
}  // namespace cuemu
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t) {
    e->t = std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
    return 0;
}
